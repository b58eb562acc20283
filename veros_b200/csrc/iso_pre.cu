// isoneutral_diffusion_pre (veros/core/isoneutral/isoneutral.py:18-229) fused with the tracer
// fluxes of isoneutral_diffusion (veros/core/isoneutral/diffusion.py:9-113) for temp and salt.
//
// One thread owns one T cell (i,j,k) and produces everything the reference stores at that index:
// the east-, north- and top-face triads (16 slopes -> Ai_ez, Ai_nz, Ai_bx, Ai_by, K_11, K_22, K_33)
// and, when FLUX is set, flux_east / flux_north / flux_top of both tracers, computed from the Ai
// values while they are still in registers (the stand-alone op re-reads 128 B/cell for each tracer).
// Threads run along the flattened (j,k) index of an x-plane, i.e. along memory: every load and
// store of a warp is contiguous; the 2x2 blocks of Ai_* leave as 16-byte stores.
//
// This kernel is bound by the FP64 pipe and by instruction issue, not by HBM (ncu: profiles/), so
// the design goal is instructions per cell:
//   * every division by a grid metric is a multiplication by a reciprocal tabulated once per call
//     (tables.cuh; the same tables hold the correctly rounded reciprocals the strict flux
//     arithmetic needs, strict.cuh);
//   * 0.5*(1+tanh(x)) = 1/(1+exp(-2x)) costs one branch-free, table-free exp (degree-12 polynomial)
//     and one Newton reciprocal seeded by rcp.approx;
//   * slope denominators use the same branch-free reciprocal; the four x- and four y-slopes of a
//     top face share two of them;
//   * masks enter as selects folded into the metric factors, never as int->double conversions.
// Work distribution: one resident wave of persistent CTAs pulls 128-cell chunks from a work counter (no partial
// last wave, next chunk's lines prefetched into L2); large grids run the horizontal and the top faces as two
// launches to halve the live state per thread; in the fused step every ninth work item of the top-face launch
// is a copy item that stages int_drhodT/S[..., tau] contiguously for the update kernel.
// Slopes/diffusivities agree with the reference to ~1e-15 of their maximum (tests/), not bit for
// bit (NumPy's SIMD tanh is not reproducible either).  The FLUXES are strict: given the stored Ai_*
// and K_* they are bit-identical to the reference's expressions (flux_device.cuh).
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "eos.cuh"
#include "flux_device.cuh"
#include "iso_pre_cell.cuh"
#include "tables.cuh"

namespace vb {

using precell::pre_cell;
using strict::Divisor;
using strict::make_divisor;


// drdT/drdS (maskT applied) for the expensive equation of state, one evaluation per cell.
__global__ void __launch_bounds__(256)
eos5_kernel(size_t ncell, int nz, const double* __restrict__ temp, const double* __restrict__ salt,
            const int32_t* __restrict__ tau_p, const uint8_t* __restrict__ maskT, const double* __restrict__ zt,
            double* __restrict__ drdT, double* __restrict__ drdS) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int tau = *tau_p;
    const int k = (int)(c % nz);
    double dT = 0.0, dS = 0.0;
    if (maskT[c]) eos_drho<5>(salt[c * 3 + tau], temp[c * 3 + tau], fabs(zt[k]), dT, dS);
    drdT[c] = dT;
    drdS[c] = dS;
}

__global__ void __launch_bounds__(256)
setup_kernel(const Grid g, const double dt, double* base, unsigned int* zero, const int nzero) {
    const int N = g.N, M = g.M, nz = g.nz;
    const Tables t = tables_at(base, N, M, nz);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    if (tid < 4) t.counters[tid] = 0u;
    for (int q = tid; q < nzero; q += nth) zero[q] = 0u;  // work queue + completion counters of the fused kernel
    for (int k = tid; k < nz; k += nth) {
        LevTab e;
        e.d4zt = make_divisor(4.0 * g.dzt[k]);
        e.dzt = make_divisor(g.dzt[k]);
        e.dzw = make_divisor(g.dzw[k]);
        e.dt_dzw = __ddiv_rn(dt, g.dzw[k]);
        e.pabs = g.zt ? fabs(g.zt[k]) : 0.0;
        const_cast<LevTab*>(t.lev)[k] = e;
    }
    for (int j = tid; j < M; j += nth) {
        RowTab e;
        const double cost = g.cost[j], dyu = g.dyu[j], cosu = g.cosu[j];
        e.dyu = make_divisor(dyu);
        e.cost = make_divisor(cost);
        e.d4ytc = make_divisor(4.0 * g.dyt[j] * cost);
        e.cdyt = make_divisor(cost * g.dyt[j]);
        e.cosu = cosu;
        e.facty = cosu * dyu;
        const_cast<RowTab*>(t.row)[j] = e;
    }
    for (int i = tid; i < N; i += nth) {
        XTab e;
        e.d4xt = make_divisor(4.0 * g.dxt[i]);
        e.dxu = g.dxu[i];
        e.pad = 0.0;
        const_cast<XTab*>(t.xt)[i] = e;
    }
    for (int q = tid; q < N * M; q += nth) {
        const int i = q / M, j = q - i * M;
        CellTab e;
        e.cdxu = make_divisor(g.cost[j] * g.dxu[i]);
        e.cdxt = make_divisor(g.cost[j] * g.dxt[i]);
        const_cast<CellTab*>(t.cell)[q] = e;
    }
}

void launch_setup_tables(cudaStream_t s, const Grid& g, double dt_tracer, double* tables, unsigned int* zero, int nzero) {
    setup_kernel<<<min(148, (g.N * g.M + 255) / 256), 256, 0, s>>>(g, dt_tracer, tables, zero, nzero);
    count_launch();
    check_launch("setup_kernel");
}

constexpr int kPreBlock = 128;

// FACES selects which faces this launch produces (bit 0 east, bit 1 north, bit 2 top).  Splitting the
// 16 slopes over two launches (east+north / top) halves the live state per thread, which raises the
// register-limited occupancy of this latency-bound kernel; the price is a second read of T and S.
template <int EOS, bool FLUX, int FACES>
__global__ void __launch_bounds__(kPreBlock, FACES == 7 ? 3 : 4)
iso_pre_kernel(const PreArgs a) {
    constexpr bool doE = (FACES & 1) != 0, doN = (FACES & 2) != 0, doT = (FACES & 4) != 0;
    const int N = a.g.N, M = a.g.M, nz = a.g.nz;
    const Tables tb = tables_at(a.tables, N, M, nz);
    const int plane_cells = M * nz;
    const int nchunks = (plane_cells + kPreBlock - 1) / kPreBlock;
    const int total = nchunks * N;
    const int tau = *a.tau;
    const double* __restrict__ T = a.temp + tau;
    const double* __restrict__ S = a.salt + tau;
    // One resident wave of persistent CTAs pulls 128-cell chunks (plane-major) from a work counter, so the
    // kernel has no partial last wave and no CTA ends up with one chunk more than another (a chunk is
    // thousands of cycles).  The counter is read two chunks ahead: the atomic's latency hides behind a
    // chunk's arithmetic, and while chunk n is computed the lines chunk n+1 will touch are pulled into L2,
    // so its loads cost an L2 hit instead of a DRAM round trip -- with ~16 resident warps per SM nothing
    // else hides that latency.
    unsigned int* counter = tb.counters + (doE ? 0 : 1);  // the two launches of a split call do not share one
    __shared__ int s_fetch[2];
    if (threadIdx.x == 0) {
        s_fetch[0] = (int)atomicAdd(counter, 1u);
        s_fetch[1] = (int)atomicAdd(counter, 1u);
    }
    __syncthreads();
    int g = s_fetch[0], gnext = s_fetch[1];
    int slot = 0;
    // With staging, every ninth work item is a copy item: 1024 cells of int_drhodT/S[..., tau] (24-byte stride)
    // into contiguous scratch.  A copy item is a few DRAM round trips of one CTA while the others compute, so the
    // copy rides on bandwidth this kernel leaves idle instead of costing the bandwidth-bound update kernel.
    constexpr int kCopyCells = 8 * kPreBlock;
    const bool staging = FLUX && doT && a.with_stage;
    const size_t ncell = (size_t)N * plane_cells;
    const int ncopy = staging ? (int)((ncell + kCopyCells - 1) / kCopyCells) : 0;
    const int items = staging ? 9 * ((total + 7) / 8) : total;
    while (g < items) {
    unsigned int fetched = 0u;
    if (threadIdx.x == 0) fetched = atomicAdd(counter, 1u);  // item after next; consumed at the end of this one
    int cg = g, cgn = gnext;
    bool is_copy = false, next_is_chunk = gnext < items;
    if (staging) {
        const int q = g / 9, r = g - 9 * q;
        is_copy = r == 8;
        cg = is_copy ? q : 8 * q + r;
        const int qn = gnext / 9, rn = gnext - 9 * qn;
        next_is_chunk = next_is_chunk && rn != 8;
        cgn = 8 * qn + rn;
    }
    if (is_copy) {
        if (cg < ncopy) {
            const size_t c0 = (size_t)cg * kCopyCells + threadIdx.x;
            double v[2][8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const size_t c = c0 + (size_t)m * kPreBlock;
                const bool in = c < ncell;
                v[0][m] = in ? __ldg(a.stage_src[0] + c * 3 + tau) : 0.0;
                v[1][m] = in ? __ldg(a.stage_src[1] + c * 3 + tau) : 0.0;
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const size_t c = c0 + (size_t)m * kPreBlock;
                if (c < ncell) {
                    a.stage[0][c] = v[0][m];
                    a.stage[1][c] = v[1][m];
                }
            }
        }
    } else if (cg < total) {
    const int i = cg / nchunks;
    const int chunk = cg - i * nchunks;
    const int p0 = chunk * kPreBlock;
    if (next_is_chunk && cgn < total) {
        const int in = cgn / nchunks;
        const int pn = (cgn - in * nchunks) * kPreBlock + (int)threadIdx.x;
        if (pn < plane_cells) {
            auto pf = [](const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); };
            const size_t pl = (size_t)plane_cells;
            const size_t cx = (size_t)in * pl + pn;
            const int i = in;  // the prefetched chunk's plane decides which neighbour planes exist
            pf(T + cx * 3); pf(S + cx * 3);
            pf(a.K_iso + cx); pf(a.maskT + cx); pf(a.maskU + cx); pf(a.maskV + cx); pf(a.maskW + cx);
            if (i + 1 < N) { pf(T + (cx + pl) * 3); pf(S + (cx + pl) * 3); pf(a.K_iso + cx + pl); pf(a.maskW + cx + pl); pf(a.maskT + cx + pl); }
            if (i >= 1) { pf(T + (cx - pl) * 3); pf(S + (cx - pl) * 3); pf(a.maskU + cx - pl); }
            if (Eos<EOS>::kExpensive) { pf(a.drdT + cx); pf(a.drdS + cx); if (i + 1 < N) { pf(a.drdT + cx + pl); pf(a.drdS + cx + pl); } }
        }
    }
    const int p = p0 + threadIdx.x;
    if (p < plane_cells) {
    pre_cell<EOS, FLUX, FACES, false>(a, tb, T, S, i, p, (size_t)i * plane_cells + p, (size_t)(i + 1) * plane_cells + p);
    }  // cell
    }  // compute chunk
    if (threadIdx.x == 0) s_fetch[slot] = (int)fetched;
    __syncthreads();
    g = gnext;
    gnext = s_fetch[slot];
    slot ^= 1;
    }  // chunk loop
}

// One resident wave: CTAs per SM of this instantiation x SMs, never more CTAs than chunks.
// Cached per (instantiation, device): the answer depends on the SM count of the device that is current.
template <typename K>
static unsigned resident_grid(K kernel, int total_chunks) {
    static std::atomic<unsigned> cache[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    unsigned full = (dev >= 0 && dev < kMaxDevices) ? cache[dev].load(std::memory_order_relaxed) : 0u;
    if (full == 0u) {
        int per_sm = 0, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPreBlock, 0) != cudaSuccess || per_sm < 1) per_sm = 3;
        full = (unsigned)(per_sm * sms);
        if (dev >= 0 && dev < kMaxDevices) cache[dev].store(full, std::memory_order_relaxed);
    }
    return (unsigned)std::max(1, std::min(total_chunks, (int)full));
}

template <int EOS, bool FLUX>
static void launch_pre_variant(cudaStream_t s, const PreArgs& a, int total_chunks) {
    // measured (profiles/): the split is +3 % on the 1 degree grid, neutral at 1 M cells, and costs one
    // launch, which small grids cannot afford.  VEROS_B200_FLAG_PRE_SINGLE / _PRE_SPLIT (descriptor) or the
    // environment variables VEROS_B200_PRE_SINGLE / VEROS_B200_PRE_SPLIT force one of the two, so that the
    // parity tests reach both instantiations on every fixture.
    bool single = (size_t)a.g.N * a.g.M * a.g.nz < 500000;
    if (a.variant == 1 || getenv("VEROS_B200_PRE_SINGLE")) single = true;
    if (a.variant == 2 || getenv("VEROS_B200_PRE_SPLIT")) single = false;
    if (single) {
        const unsigned full = resident_grid(iso_pre_kernel<EOS, FLUX, 7>, 1 << 30);
        iso_pre_kernel<EOS, FLUX, 7><<<std::min(full, (unsigned)total_chunks), kPreBlock, 0, s>>>(a);
        count_launch();
    } else {
        const unsigned full3 = resident_grid(iso_pre_kernel<EOS, FLUX, 3>, 1 << 30);
        const unsigned full4 = resident_grid(iso_pre_kernel<EOS, FLUX, 4>, 1 << 30);
        iso_pre_kernel<EOS, FLUX, 3><<<std::min(full3, (unsigned)total_chunks), kPreBlock, 0, s>>>(a);
        iso_pre_kernel<EOS, FLUX, 4><<<std::min(full4, (unsigned)total_chunks), kPreBlock, 0, s>>>(a);
        count_launch(2);
    }
}

template <int EOS>
static void launch_pre_eos(cudaStream_t s, const PreArgs& a, int grid) {
    if (a.with_flux)
        launch_pre_variant<EOS, true>(s, a, grid);
    else
        launch_pre_variant<EOS, false>(s, a, grid);
}

void launch_iso_pre(cudaStream_t s, const PreArgs& a0, bool profile) {
    PreArgs a = a0;
    const int N = a.g.N, M = a.g.M, nz = a.g.nz;
    const size_t ncell = (size_t)N * M * nz;
    if (ncell == 0) return;
    a.two_rd = 2.0 / a.iso_dslope;
    a.m2c0 = -2.0 * a.iso_slopec / a.iso_dslope;
    a.s_max = (345.0 + a.iso_slopec / a.iso_dslope) * a.iso_dslope;
    if (!a.tables_ready) launch_setup_tables(s, a.g, a.dt_tracer, a.tables);
    if (a.eos == 5) {
        eos5_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, s>>>(ncell, nz, a.temp, a.salt, a.tau, a.maskT, a.g.zt,
                                                                   a.drdT, a.drdS);
        count_launch();
        if (!check_launch("eos5_kernel")) return;
    }
    const int nchunks = (M * nz + kPreBlock - 1) / kPreBlock;
    const int grid = nchunks * N;  // chunks in total; the launch is one resident wave pulling them
    if (profile) prof_mark(s, 1);
    switch (a.eos) {
    case 1: launch_pre_eos<1>(s, a, grid); break;
    case 2: launch_pre_eos<2>(s, a, grid); break;
    case 3: launch_pre_eos<3>(s, a, grid); break;
    case 4: launch_pre_eos<4>(s, a, grid); break;
    default: launch_pre_eos<5>(s, a, grid); break;
    }
    if (profile) prof_mark(s, 2);
    check_launch("iso_pre_kernel");
}

}  // namespace vb
