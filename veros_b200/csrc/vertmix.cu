// vertmix_tempsalt (veros/core/thermodynamics.py:248-288): implicit vertical mixing of temperature and
// salinity with kappaH and the surface fluxes -- the step right after the isoneutral path
// (SURVEY.md section 8f, rank 1).  One kernel: the reference's five coefficient arrays, two
// solve_implicit calls and four whole-array tendency passes become
//   phase B  per cell: matrix rows from kappaH (both tracers share them), right-hand sides
//   phase C  per column: LAPACK dgtsv replay on two right-hand sides (tdma_device.cuh)
//   phase D  per cell: where(water, sol, tr), d{temp,salt}_vmix = (new - old) / dt_tracer
// on tiles of whole columns staged in shared memory.  Every operation is an explicitly rounded
// intrinsic in the reference's order, so the result equals the NumPy backend's bit for bit.
// The boundary treatment that follows in the reference (enforce_boundaries, :290-297) is the halo
// exchange of the host layer (csrc/halo.cu, veros_b200/decomp.py).
#include "common.cuh"
#include "strict.cuh"
#include "tdma_device.cuh"

namespace vb {
namespace {

using strict::add;
using strict::Divisor;
using strict::make_divisor;
using strict::mul;
using strict::sub;

constexpr int kVmixBlock = 128;

__global__ void __launch_bounds__(kVmixBlock, 6)
vertmix_kernel(const VmixArgs a, const int cols, const int pitch) {
    extern __shared__ double sm[];
    const int N = a.N, M = a.M, nz = a.nz;
    const int i = blockIdx.y;
    const int j0 = blockIdx.x * cols;
    const int ncols = min(cols, M - j0);
    const int ncells = ncols * nz;
    const int taup1 = *a.taup1;
    const double dt = a.dt_tracer;
    const Divisor ddt = make_divisor(dt);
    const size_t base = ((size_t)i * M + j0) * nz;
    const bool i_int = i >= 2 && i < N - 2;

    Divisor* ddzt = reinterpret_cast<Divisor*>(sm);  // dzt[k]
    double* dt_dzw = reinterpret_cast<double*>(ddzt + nz);  // dt_tracer / dzw[k]
    const int tile = cols * pitch;
    double* L = dt_dzw + nz;
    double* D = L + tile;
    double* U = D + tile;
    double* R0 = U + tile;
    double* R1 = R0 + tile;
    int* ksv = reinterpret_cast<int*>(R1 + tile);

    for (int k = threadIdx.x; k < nz; k += kVmixBlock) {
        ddzt[k] = make_divisor(a.dzt[k]);
        dt_dzw[k] = strict::div(dt, a.dzw[k]);
    }
    for (int q = threadIdx.x; q < ncols; q += kVmixBlock) ksv[q] = a.kbot[i * M + j0 + q] - 1;
    __syncthreads();

    // ---- phase B: matrix and right-hand sides of the interior columns ------------------------------
    if (i_int) {
        for (int idx = threadIdx.x; idx < ncells; idx += kVmixBlock) {
            const int q = idx / nz, k = idx - q * nz;
            const int j = j0 + q;
            if (j < 2 || j >= M - 2) continue;
            const size_t c = base + idx;
            const int s = q * pitch + k;
            const double kap = (k < nz - 1) ? __ldg(a.kappaH + c) : 0.0;
            const double kapm = (k > 0) ? __ldg(a.kappaH + c - 1) : 0.0;
            double r0 = a.temp[c * 3 + taup1], r1 = a.salt[c * 3 + taup1];
            if (k == nz - 1) {  // surface fluxes, thermodynamics.py:276,282
                const size_t c2 = (size_t)i * M + j;
                r0 = add(r0, strict::div(mul(dt, __ldg(a.forc_temp + c2)), ddzt[k]));
                r1 = add(r1, strict::div(mul(dt, __ldg(a.forc_salt + c2)), ddzt[k]));
            }
            const int ks = ksv[q];
            const double del = (k < nz - 1) ? mul(dt_dzw[k], kap) : 0.0;       // :267-270
            const double delm = (k > 0) ? mul(dt_dzw[k - 1], kapm) : 0.0;
            // b_tri (:272) on rows k >= 1, b_tri_edge (:273) on the bottom water row
            D[s] = (k == ks) ? add(1.0, strict::div(del, ddzt[k])) : add(1.0, strict::div(add(del, delm), ddzt[k]));
            U[s] = (k < nz - 1) ? strict::div(-del, ddzt[k]) : 0.0;             // c_tri, :274
            if (k > 0) L[s - 1] = (k > ks) ? strict::div(-delm, ddzt[k]) : 0.0;  // a_tri, :271; 0 on the edge row
            R0[s] = r0;
            R1[s] = r1;
        }
    }
    __syncthreads();

    // ---- phase C: one thread per water column ----------------------------------------------------------
    if (i_int) {
        for (int q = threadIdx.x; q < ncols; q += kVmixBlock) {
            const int j = j0 + q;
            const int ks = ksv[q];
            if (j >= 2 && j < M - 2 && ks >= 0) {
                const int o = q * pitch;
                dgtsv_column<2>(ks, nz, L + o, D + o, U + o, R0 + o, R1 + o);
            }
        }
    }
    __syncthreads();

    // ---- phase D: every cell of the array (ghost cells get (tr - tr) / dt like the reference) ---------
    for (int idx = threadIdx.x; idx < ncells; idx += kVmixBlock) {
        const int q = idx / nz, k = idx - q * nz;
        const int j = j0 + q;
        const size_t c = base + idx;
        const int s = q * pitch + k;
        const int ks = ksv[q];
        const bool solved = i_int && j >= 2 && j < M - 2 && ks >= 0 && k >= ks;
        const double old0 = a.temp[c * 3 + taup1], old1 = a.salt[c * 3 + taup1];
        const double nw0 = solved ? R0[s] : old0, nw1 = solved ? R1[s] : old1;
        if (solved) {
            a.temp[c * 3 + taup1] = nw0;
            a.salt[c * 3 + taup1] = nw1;
        }
        a.dtemp_vmix[c] = strict::div(sub(nw0, old0), ddt);  // :287-288
        a.dsalt_vmix[c] = strict::div(sub(nw1, old1), ddt);
    }
}

}  // namespace

void launch_vertmix(cudaStream_t s, const VmixArgs& a) {
    const int N = a.N, M = a.M, nz = a.nz;
    if (N <= 0 || M <= 0 || nz <= 0) return;
    const int pitch = (nz + 1) | 1;  // odd (no bank conflicts) and > nz: the element past a column belongs to nobody
    int cols = max(1, 640 / nz);
    cols = min(cols, M);
    const int want_tiles = 4 * 148;  // small grids: spread over the SMs
    if (((M + cols - 1) / cols) * N < want_tiles) {
        const int per_row = (want_tiles + N - 1) / N;
        cols = max(1, (M + per_row - 1) / per_row);
    }
    const size_t smem = (size_t)nz * (sizeof(Divisor) + 8) + 8 * ((size_t)5 * cols * pitch + (cols + 1) / 2 + 1);
    allow_big_smem(vertmix_kernel, 200 * 1024);  // per device, not per process
    dim3 grid((M + cols - 1) / cols, N);
    vertmix_kernel<<<grid, kVmixBlock, smem, s>>>(a, cols, pitch);
    count_launch();
    check_launch("vertmix_kernel");
}

}  // namespace vb
