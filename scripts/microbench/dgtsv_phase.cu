// Microbenchmark: the column-solve phase of update_kernel / vertmix_kernel in isolation.
// CTAs of 128 threads hold a tile of `cols` columns x nz levels (5 arrays, odd pitch) in shared memory;
// all threads refill the tile with a diagonally dominant system, `cols` threads of warp 0 run
// dgtsv_column<2>, repeat.  Reports cycles per level of the solve alone (fill time subtracted by a
// second run without the solve) as a function of CTAs per SM and columns per CTA.
// Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I veros_b200/csrc -o scripts/microbench/dgtsv_phase scripts/microbench/dgtsv_phase.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "tdma_device.cuh"

using namespace vb;

__global__ void __launch_bounds__(128, 6) phase(double* out, int nz, int cols, int iters, int solve, int rotate) {
    extern __shared__ double sm[];
    const int pitch = (nz + 1) | 1;
    const int tile = cols * pitch;
    double* L = sm + 1;
    double* D = L + tile;
    double* U = D + tile;
    double* R0 = U + tile;
    double* R1 = R0 + tile;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        for (int idx = threadIdx.x; idx < cols * nz; idx += blockDim.x) {
            const int q = idx / nz, k = idx - q * nz;
            const int s = q * pitch + k;
            const double del = 0.3 + 1e-3 * ((idx + it + blockIdx.x) % 97), delm = 0.2 + 1e-3 * ((idx + 2 * it) % 89);
            D[s] = 1.0 + del + delm;
            U[s] = k < nz - 1 ? -del : 0.0;
            if (k > 0) L[s - 1] = -delm;
            R0[s] = 10.0 + 1e-2 * (idx % 13);
            R1[s] = 35.0 + 1e-3 * (idx % 7);
        }
        __syncthreads();
        const int vt = (threadIdx.x + 128 - (rotate ? 32 * ((blockIdx.x / 148) & 3) : 0)) & 127;
        if (solve && vt < cols) {
            const int o = vt * pitch;
#ifdef OLD_DGTSV
            dgtsv_column<2>(0, nz, 1, L + o, D + o, U + o, R0 + o, R1 + o);
#else
            dgtsv_column<2>(0, nz, L + o, D + o, U + o, R0 + o, R1 + o);
#endif
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < cols * nz; idx += blockDim.x) {
            const int q = idx / nz, k = idx - q * nz;
            acc += R0[q * pitch + k] + R1[q * pitch + k];
        }
        __syncthreads();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    const int nz = argc > 1 ? atoi(argv[1]) : 50;
    double* out;
    cudaMalloc(&out, 148 * 8 * 128 * 8);
    cudaFuncSetAttribute(phase, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int mhz;
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const int iters = 200;
    const int rotate = argc > 2 ? atoi(argv[2]) : 0;  // 1: the solver warp rotates with the CTA's launch position
    for (int ctas = 1; ctas <= 6; ctas += (ctas < 2 ? 1 : 2)) {
        for (int cols : {4, 12, 24, 32}) {
            const size_t smem = 8 * ((size_t)5 * cols * ((nz + 1) | 1) + 4);
            if (smem * ctas > 220 * 1024) continue;
            float ms[2];
            for (int solve = 0; solve < 2; ++solve) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                phase<<<148 * ctas, 128, smem>>>(out, nz, cols, 5, solve, rotate);
                cudaEventRecord(e0);
                phase<<<148 * ctas, 128, smem>>>(out, nz, cols, iters, solve, rotate);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[solve], e0, e1);
            }
            const double cyc = (ms[1] - ms[0]) * 1e-3 * mhz * 1e3 / ((double)iters * nz);
            printf("nz %3d  CTAs/SM %d  cols/CTA %2d : %7.1f cycles per level (solve %.3f ms, fill-only %.3f ms per %d tiles)\n", nz,
                   ctas, cols, cyc, ms[1], ms[0], iters);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
