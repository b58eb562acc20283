// Microbenchmark: FP64 FMA dependent-issue latency and throughput on the target GPU, as a function of
// independent chains per thread (ILP) and resident warps per SM.  Build: nvcc -arch=sm_100a -O3 -o dfma dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chains(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int q = 0; q < ILP; ++q) x[q] = threadIdx.x * 1e-3 + q;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) x[q] = fma(x[q], a, b);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < ILP; ++q) s += x[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int warps_per_sm, double* out) {
    const int iters = 20000;
    int sms = 148;
    dim3 grid(sms), block(32 * warps_per_sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    chains<ILP><<<grid, block>>>(out, 100, 0.999, 1e-3);
    cudaEventRecord(e0);
    chains<ILP><<<grid, block>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int mhz;
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * mhz * 1e3;
    const double per_warp_instr = (double)iters * ILP;
    printf("ILP %2d warps/SM %2d: %.2f cycles per DFMA per warp-chain step (latency view), %.3f DFMA warp-instr/clk/SM\n", ILP,
           warps_per_sm, cycles / iters, per_warp_instr * warps_per_sm / cycles);
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 1024 * 8);
    run<1>(1, out); run<1>(4, out);
    run<2>(4, out); run<4>(4, out); run<8>(4, out); run<16>(4, out);
    run<1>(12, out); run<2>(12, out); run<4>(12, out);
    run<1>(16, out); run<2>(16, out); run<1>(32, out); run<1>(64 / 2, out);
    cudaFree(out);
    return 0;
}
