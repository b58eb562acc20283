// Column (tridiagonal) solves.
//
//  * solve_implicit_kernel: veros.core.utilities.solve_implicit + solve_tridiagonal on the model's
//    native (X,Y,nz) z-contiguous layout (utilities.py:51-59, operators.py:60-77, tdma_.py:59-68).
//    A CTA stages a tile of whole columns in shared memory with coalesced, streaming loads, one
//    thread per column then runs dgtsv's elimination out of shared memory (odd pitch -> no bank
//    conflicts), and the tile is written back coalesced.  No layout conversion in HBM: 42 B/cell.
//  * tdma_zmajor_kernel: drop-in for the reference's TridiagKernel (cuda_tdma_kernels.cu:19-70):
//    same buffers, same z-major layout, same Thomas recurrence; one thread per system.
#include "common.cuh"
#include "tdma_device.cuh"

namespace vb {

// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
solve_implicit_kernel(int ncol, int nz, int cols_per_tile, int pitch, const double* __restrict__ a,
                      const double* __restrict__ b, const double* __restrict__ c, const double* __restrict__ d,
                      const uint8_t* __restrict__ water, const uint8_t* __restrict__ edge,
                      const double* __restrict__ b_edge, const double* __restrict__ d_edge,
                      double* __restrict__ out) {
    extern __shared__ double smem[];
    const int tile_elems = cols_per_tile * pitch;
    double* L = smem + 1;  // dgtsv_column may touch one element before / after each column (never uses it)
    double* D = L + tile_elems;
    double* U = D + tile_elems;
    double* R = U + tile_elems;
    int* kfirst = reinterpret_cast<int*>(R + tile_elems);  // first water row of each column

    const int col0 = blockIdx.x * cols_per_tile;
    const int ncols = min(cols_per_tile, ncol - col0);
    const int ncells = ncols * nz;
    const size_t base = (size_t)col0 * nz;

    for (int c_ = threadIdx.x; c_ < ncols; c_ += blockDim.x) kfirst[c_] = nz;
    __syncthreads();

    for (int idx = threadIdx.x; idx < ncells; idx += blockDim.x) {
        const int col = idx / nz, k = idx - col * nz;
        const size_t g = base + idx;
        const bool w = water[g] != 0, e = edge[g] != 0;
        const int s = col * pitch + k;
        double bb = b[g], dd = d[g];
        if (e && b_edge) bb = b_edge[g];
        if (e && d_edge) dd = d_edge[g];
        D[s] = w ? bb : 1.0;
        R[s] = w ? dd : 0.0;
        U[s] = (w && k < nz - 1) ? c[g] : 0.0;
        if (k >= 1) L[s - 1] = (w && !e) ? a[g] : 0.0;
        if (w) atomicMin(&kfirst[col], k);
    }
    __syncthreads();

    for (int col = threadIdx.x; col < ncols; col += blockDim.x) {
        const int o = col * pitch;
        // rows below the first water row are identity rows with zero rhs: they do not change any
        // later row bitwise (fact * 0 terms), so the elimination starts at the first water row
        dgtsv_column<1>(kfirst[col], nz, L + o, D + o, U + o, R + o, nullptr);
    }
    __syncthreads();

    for (int idx = threadIdx.x; idx < ncells; idx += blockDim.x) {
        const int col = idx / nz, k = idx - col * nz;
        const size_t g = base + idx;
        out[g] = water[g] ? R[col * pitch + k] : 0.0;
    }
}

void launch_solve_implicit(cudaStream_t s, int ncol, int nz, const double* a, const double* b, const double* c,
                           const double* d, const uint8_t* water, const uint8_t* edge, const double* b_edge,
                           const double* d_edge, double* out) {
    if (ncol <= 0 || nz <= 0) return;
    const int pitch = (nz + 1) | 1;  // odd (no bank conflicts) and > nz: the element past a column belongs to nobody
    int cols = (56 * 1024) / (4 * 8 * pitch);
    cols = max(1, min(cols, 64));
    // small problems: keep at least ~2 CTAs per SM busy
    const int want_tiles = 2 * 148;
    if ((ncol + cols - 1) / cols < want_tiles) cols = max(1, (ncol + want_tiles - 1) / want_tiles);
    const size_t smem = (size_t)4 * 8 * cols * pitch + sizeof(int) * cols + 24;  // + the two guard elements
    allow_big_smem(solve_implicit_kernel, 200 * 1024);  // per device, not per process
    const int grid = (ncol + cols - 1) / cols;
    solve_implicit_kernel<<<grid, 256, smem, s>>>(ncol, nz, cols, pitch, a, b, c, d, water, edge, b_edge, d_edge, out);
    count_launch();
    check_launch("solve_implicit_kernel");
}

// -------------------------------------------------------------------------------------------------
// z-major Thomas kernel with the reference's buffer convention: a,b,c,d,(out),(workspace) are
// [depth][nsys]; masks were applied by the caller (tdma_.py:63-66).
template <typename T>
__global__ void __launch_bounds__(128)
tdma_zmajor_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                   const T* __restrict__ d, T* __restrict__ cp, T* __restrict__ dp, int nsys, int depth) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nsys; idx += blockDim.x * gridDim.x) {
        size_t j = idx;
        const T b0 = b[j];
        T cm1 = c[j] / b0;
        T dm1 = d[j] / b0;
        cp[j] = cm1;
        dp[j] = dm1;
        for (int k = 1; k < depth; ++k) {
            j += nsys;
            const T ai = a[j];
            const T denom = T(1) / (b[j] - ai * cm1);
            cm1 = c[j] * denom;
            dm1 = (d[j] - ai * dm1) * denom;
            cp[j] = cm1;
            dp[j] = dm1;
        }
        T x = dm1;
        for (int k = depth - 2; k >= 0; --k) {
            j -= nsys;
            x = dp[j] - cp[j] * x;
            dp[j] = x;
        }
    }
}

template <typename T>
static void launch_zmajor(cudaStream_t s, int nsys, int depth, const T* a, const T* b, const T* c, const T* d,
                          T* out, T* work) {
    if (nsys <= 0 || depth <= 0) return;
    const int block = 128;
    const int grid = (nsys + block - 1) / block;
    tdma_zmajor_kernel<T><<<grid, block, 0, s>>>(a, b, c, d, work, out, nsys, depth);
    count_launch();
    check_launch("tdma_zmajor_kernel");
}

void launch_tdma_zmajor_f64(cudaStream_t s, int nsys, int depth, const double* a, const double* b, const double* c,
                            const double* d, double* out, double* work) {
    launch_zmajor<double>(s, nsys, depth, a, b, c, d, out, work);
}
void launch_tdma_zmajor_f32(cudaStream_t s, int nsys, int depth, const float* a, const float* b, const float* c,
                            const float* d, float* out, float* work) {
    launch_zmajor<float>(s, nsys, depth, a, b, c, d, out, work);
}

}  // namespace vb
