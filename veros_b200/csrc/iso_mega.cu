// The fused isoneutral step as ONE persistent kernel (veros/core/thermodynamics.py:430-432: isoneutral_diffusion_pre,
// isoneutral_diffusion(temp), isoneutral_diffusion(salt) incl. the implicit column solves).
//
// Why one kernel.  As separate launches the step is a chain of phases that each stress ONE resource while the others
// idle: the TEOS-10 pass and the tracer update stream memory, the slope kernel is bound by FP64 issue and latency, the
// column solves are serial latency chains run by a few lanes per CTA (profiles/: DRAM 26-50 % busy, FP64 pipe 40 %,
// 22 % of the update kernel's stall samples at the barrier behind the solves).  And every phase hands its results to
// the next through HBM: the six flux arrays, the TEOS-10 derivatives, the staged int_drhod* copies and the T-point
// dissipation are written and re-read (4.9 GB of DRAM traffic per 1 degree step against 1.8 GB algorithmic).
//
// Here all phases are WORK ITEMS of one launch.  The step is cut along x into planes; per plane there are
//     E(i)  TEOS-10 derivatives of plane i              (eq_of_state_type 5 only)
//     C(i)  contiguous copy of int_drhodT/S[i,...,tau]  (enable_conserve_energy only)
//     P(i)  slopes, mixing tensor and fluxes of plane i (needs E(i), E(i+1))
//     U(i)  divergence, column solve, tendencies, dissipation of plane i (needs P(i-1), P(i), C(i-1..i+1))
// items, queued round by round as E(r), C(r), P(r-lag), U(r-2 lag), the lag chosen so that dependent items are a
// full wave of CTAs apart in the queue (a blocked item idles its SM slot).  One resident wave of CTAs pulls items from an atomic
// counter; an item waits (thread 0 polls a per-plane completion counter) until the items it depends on have finished.
// Dependencies only point to EARLIER queue positions and an item is owned by a CTA that is already running, so the
// scheme cannot deadlock and needs no co-residency guarantee (no cooperative launch, capturable in a CUDA graph).
// What this buys:
//   * at any moment an SM holds CTAs in different phases -- FP64-heavy slope items overlap memory-streaming update
//     items and the serial column solves of a third CTA (the "de-phasing" the lock-step kernels cannot do);
//   * all inter-phase scratch lives in a RING of a few x-planes (fluxes, derivatives, staged copies, dissipation:
//     96 B/cell x ring planes, 10-35 MB) that is produced and consumed within microseconds and therefore stays in
//     the 126 MB L2: it never reaches HBM.  Ring slots are recycled under the same counters (write-after-read);
//   * the update tiles are as wide as shared memory allows (up to 32 columns = one full solver warp) because the
//     serial solve of one CTA no longer idles the SM;
//   * one launch per step (+ the table setup) instead of four or five: the small-grid latency path.
// Memory ordering: producers publish with red.release.gpu (MEMBAR.GPU, no L1 invalidation), consumers poll with
// ld.relaxed.gpu and read everything another CTA of this launch produced with ld.global.cg (served by L2, the point of
// coherence), so no CTA ever flushes its SM's L1.  (__threadfence / ld.acquire would emit CCTL.IVALL: checked in SASS.)
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "eos.cuh"
#include "iso_pre_cell.cuh"
#include "iso_update_phases.cuh"
#include "tables.cuh"

namespace vb {

using namespace upd;

constexpr int kMegaBlock = 128;
constexpr int kMegaPreCells = 256;    // cells per P item
constexpr int kMegaCopyCells = 1024;  // cells per E / C item

struct MegaArgs {
    PreArgs p;      // scratch pointers (flux, drdT/S, stage) are ring arrays
    DiffArgs d;     // stage_x = ring arrays
    Scratch f;      // fe / fn / ft / diss ring arrays
    int ring;       // x-planes in the scratch ring
    int cols, pitch;            // update tile: columns, shared-memory pitch
    int nE, nC, nP, nU;         // items per plane
    int lagP, lagU;             // queue order: round r holds E(r), C(r), P(r - lagP), U(r - lagU)
    unsigned int* sync;         // [0] queue head, [1] unused, then eosDone[N], copyDone[N], preDone[N], updDone[N]
    unsigned long long* stats;  // [type]: ns spent waiting, [5 + type]: ns spent working, [10 + type]: items (thread 0's clock)
    double fac_diss, gr;
};

namespace {

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void add_release(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// thread 0 only: block until *p >= target (all producers of one plane have published)
__device__ __forceinline__ void wait_for(const unsigned* p, unsigned target) {
    if (ld_relaxed(p) >= target) return;
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (ld_relaxed(p) < target) {
        __nanosleep(40);
        if ((++spins & 0xfffu) == 0u && global_ns() - t0 > 10000000000ull) __trap();  // 10 s: a bug, not a wait
    }
}

enum ItemType { kNone = 0, kEos, kCopy, kPre, kUpd };

struct Item {
    int type, plane, idx;
};

}  // namespace

template <int EOS, bool ENERGY>
__global__ void __launch_bounds__(kMegaBlock, 3)
iso_mega_kernel(const MegaArgs m) {
    extern __shared__ double sm[];
    __shared__ int s_next;
    const int N = m.p.g.N, M = m.p.g.M, nz = m.p.g.nz, R = m.ring;
    const int plane_cells = M * nz;
    const size_t plane = (size_t)plane_cells;
    const Tables tb = tables_at(m.p.tables, N, M, nz);
    const int tau = *m.p.tau;
    const double* __restrict__ T = m.p.temp + tau;
    const double* __restrict__ S = m.p.salt + tau;
    unsigned* const queue = m.sync;
    unsigned* const eosDone = m.sync + 2;
    unsigned* const copyDone = eosDone + N;
    unsigned* const preDone = copyDone + N;
    unsigned* const updDone = preDone + N;
    const int per_round = m.nE + m.nC + m.nP + m.nU;
    const int total = per_round * (N + m.lagU);

    // update-side constants and per-level tables (once per CTA)
    const UpdConst u = upd_const(m.d, m.pitch, m.fac_diss, m.gr);
    LevelTabs lv;
    lv.ddzt = reinterpret_cast<Divisor*>(sm);
    lv.ddzw = lv.ddzt + nz;
    lv.dt_dzw = reinterpret_cast<double*>(lv.ddzw + nz);
    const TileBuf b = tile_buf_at(lv.dt_dzw + nz, m.cols, m.pitch, 2);
    fill_level_tabs(m.d, lv, nz, u.dt, threadIdx.x, kMegaBlock);

    auto decode = [&](int g) {
        Item it;
        const int r = g / per_round;
        int w = g - r * per_round;
        it.type = kNone;
        it.plane = 0;
        it.idx = 0;
        if (w < m.nE) {
            if (r < N) { it.type = kEos; it.plane = r; it.idx = w; }
            return it;
        }
        w -= m.nE;
        if (w < m.nC) {
            if (r < N) { it.type = kCopy; it.plane = r; it.idx = w; }
            return it;
        }
        w -= m.nC;
        if (w < m.nP) {
            if (r >= m.lagP && r - m.lagP < N) { it.type = kPre; it.plane = r - m.lagP; it.idx = w; }
            return it;
        }
        w -= m.nP;
        if (r - m.lagU >= 1 && r - m.lagU <= N - 2) { it.type = kUpd; it.plane = r - m.lagU; it.idx = w; }
        return it;
    };

    // thread 0: everything item `it` depends on has been published (read-after-write), and every reader of the
    // ring slots it is about to overwrite has finished (write-after-read)
    auto wait_deps = [&](const Item& it) {
        const int i = it.plane;
        switch (it.type) {
        case kEos:  // drd slot of plane i-R was read by P(i-R) and P(i-R-1)
            if (i - R >= 0) wait_for(preDone + i - R, m.nP);
            if (i - R - 1 >= 0) wait_for(preDone + i - R - 1, m.nP);
            break;
        case kCopy:  // stage slot of plane i-R was read by U(i-R-1), U(i-R), U(i-R+1)
            for (int q = i - R - 1; q <= i - R + 1; ++q)
                if (q >= 1 && q <= N - 2) wait_for(updDone + q, m.nU);
            break;
        case kPre:
            if (m.nE) {
                wait_for(eosDone + i, m.nE);
                if (i + 1 < N) wait_for(eosDone + i + 1, m.nE);
            }
            for (int q = i - R; q <= i - R + 1; ++q)  // flux slot of plane i-R was read by U(i-R), U(i-R+1)
                if (q >= 1 && q <= N - 2) wait_for(updDone + q, m.nU);
            break;
        case kUpd:
            wait_for(preDone + i - 1, m.nP);
            wait_for(preDone + i, m.nP);
            if (m.nC) {
                wait_for(copyDone + i - 1, m.nC);
                wait_for(copyDone + i, m.nC);
                wait_for(copyDone + i + 1, m.nC);
            }
            if (i - R >= 1) wait_for(updDone + i - R, m.nU);  // dissipation slot of plane i-R
            break;
        default:
            break;
        }
    };

    unsigned long long t_start = 0;  // thread 0: when the current item's work began
    if (threadIdx.x == 0) {
        const int g = (int)atomicAdd(queue, 1u);
        const unsigned long long t0 = global_ns();
        if (g < total) {
            const Item nx = decode(g);
            wait_deps(nx);
            t_start = global_ns();
            atomicAdd(m.stats + nx.type, t_start - t0);
        }
        s_next = g;
    }
    __syncthreads();
    int g = s_next;

    while (g < total) {
        const Item it = decode(g);
        const int i = it.plane;
        const size_t ring_c = (size_t)(i % R) * plane;
        if (it.type == kEos) {
            // ---- E(i): drdT, drdS (maskT applied) of plane i, one evaluation per cell -------------------------
            const int p0 = it.idx * kMegaCopyCells + threadIdx.x;
#pragma unroll 2
            for (int q = 0; q < kMegaCopyCells / kMegaBlock; ++q) {
                const int p = p0 + q * kMegaBlock;
                if (p < plane_cells) {
                    const size_t c = (size_t)i * plane + p;
                    const int k = p % nz;
                    double dT = 0.0, dS = 0.0;
                    if (m.p.maskT[c]) eos_drho<5>(__ldg(S + c * 3), __ldg(T + c * 3), __ldg(&tb.lev[k].pabs), dT, dS);
                    m.p.drdT[ring_c + p] = dT;
                    m.p.drdS[ring_c + p] = dS;
                }
            }
        } else if (it.type == kCopy) {
            // ---- C(i): int_drhodT/S[i, ..., tau] (24-byte stride) -> contiguous ring plane --------------------
            const int p0 = it.idx * kMegaCopyCells + threadIdx.x;
            double v[2][kMegaCopyCells / kMegaBlock];
#pragma unroll
            for (int q = 0; q < kMegaCopyCells / kMegaBlock; ++q) {
                const int p = p0 + q * kMegaBlock;
                const size_t c = (size_t)i * plane + p;
                const bool in = p < plane_cells;
                v[0][q] = in ? __ldg(m.p.stage_src[0] + c * 3 + tau) : 0.0;
                v[1][q] = in ? __ldg(m.p.stage_src[1] + c * 3 + tau) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < kMegaCopyCells / kMegaBlock; ++q) {
                const int p = p0 + q * kMegaBlock;
                if (p < plane_cells) {
                    m.p.stage[0][ring_c + p] = v[0][q];
                    m.p.stage[1][ring_c + p] = v[1][q];
                }
            }
        } else if (it.type == kPre) {
            // ---- P(i): slopes, mixing tensor, fluxes ---------------------------------------------------------
            const size_t ring_e = (size_t)((i + 1) % R) * plane;
            const int p0 = it.idx * kMegaPreCells + threadIdx.x;
            for (int q = 0; q < kMegaPreCells / kMegaBlock; ++q) {
                const int p = p0 + q * kMegaBlock;
                if (p < plane_cells) precell::pre_cell<EOS, true, 7, true>(m.p, tb, T, S, i, p, ring_c + p, ring_e + p);
            }
        } else if (it.type == kUpd) {
            // ---- U(i): one tile of whole columns -------------------------------------------------------------
            const bool skip = (i == 1 && m.d.skip_west_ring) || (i == N - 2 && m.d.skip_east_ring);
            if (!skip) {
                TileGeom tg = tile_geom(u, i, 1 + it.idx * m.cols, m.cols);
                const size_t in_plane = (size_t)tg.j0 * nz;
                tg.sb_c = ring_c + in_plane;
                tg.sb_w = (size_t)((i - 1) % R) * plane + in_plane;
                tg.sb_e = (size_t)((i + 1) % R) * plane + in_plane;
                tg.db = tg.sb_c;
                fill_tile_tabs(m.d, u, b, tg, threadIdx.x, kMegaBlock);
                __syncthreads();
                phase_b<2, false, ENERGY, true>(m.d, m.f, u, lv, b, tg, threadIdx.x, kMegaBlock);
                __syncthreads();
                phase_c<2>(u, b, tg, threadIdx.x, kMegaBlock);
                __syncthreads();
                phase_d<2, false, ENERGY, true>(m.d, m.f, u, lv, b, tg, threadIdx.x, kMegaBlock);
            }
        }
        __syncthreads();  // every store of this item has been issued (and shared memory may be reused)
        if (threadIdx.x == 0) {
            unsigned* done = it.type == kEos ? eosDone : it.type == kCopy ? copyDone : it.type == kPre ? preDone : updDone;
            if (it.type != kNone) add_release(done + i);
            const unsigned long long t0 = global_ns();
            atomicAdd(m.stats + 5 + it.type, t0 - t_start);
            atomicAdd(m.stats + 10 + it.type, 1ull);
            const int gn = (int)atomicAdd(queue, 1u);
            t_start = t0;
            if (gn < total) {
                const Item nx = decode(gn);
                wait_deps(nx);
                t_start = global_ns();
                atomicAdd(m.stats + nx.type, t_start - t0);
            }
            s_next = gn;
        }
        __syncthreads();
        g = s_next;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
namespace {

struct MegaPlan {
    int ring, cols, pitch, nE, nC, nP, nU, lagP, lagU;
    size_t smem;
};

MegaPlan mega_plan(int N, int M, int nz, int eos, int energy) {
    MegaPlan q;
    const int plane_cells = M * nz;
    q.pitch = (nz + 1) | 1;
    // update tile: as many whole columns as fit into ~1/3 of an SM's shared memory, at most one solver warp's worth
    const size_t lev_bytes = (size_t)nz * (2 * sizeof(strict::Divisor) + 8);
    const size_t budget = 72 * 1024;
    const size_t per_col = (size_t)5 * 8 * q.pitch + 2 * sizeof(strict::Divisor) + 4;
    int cols = (int)((budget - std::min(budget, lev_bytes + 64)) / per_col);
    cols = std::max(1, std::min(cols, 32));
    cols = std::min(cols, M - 2);
    // small grids: enough update items per plane to spread over the SMs
    const int want_items = 4 * 148;
    if (((M - 2 + cols - 1) / cols) * (N - 2) < want_items) {
        const int per_plane = (want_items + (N - 2) - 1) / (N - 2);
        cols = std::max(1, std::min(cols, (M - 2 + per_plane - 1) / per_plane));
    }
    q.cols = cols;
    q.smem = lev_bytes + 8 * ((size_t)5 * cols * q.pitch + (size_t)cols * 4 + (cols + 1) / 2 + 1) + 16;
    q.nE = eos == 5 ? (plane_cells + kMegaCopyCells - 1) / kMegaCopyCells : 0;
    q.nC = energy ? (plane_cells + kMegaCopyCells - 1) / kMegaCopyCells : 0;
    q.nP = (plane_cells + kMegaPreCells - 1) / kMegaPreCells;
    q.nU = (M - 2 + cols - 1) / cols;
    // Queue order.  A CTA that pulls an item whose producers are still running blocks (and idles its SM slot), so
    // dependent items are queued `lag` rounds apart, where one lag covers more items than there are resident CTAs
    // (<= 4 per SM on <= 160 SMs): by the time U(i) is pulled, every P(i) item was pulled a full wave earlier.
    const int per_round = q.nE + q.nC + q.nP + q.nU;
    int lag = 1 + (640 + per_round - 1) / per_round;
    if (const char* e = getenv("VEROS_B200_MEGA_LAG")) lag = std::max(1, atoi(e));  // tuning knob
    lag = std::min(lag, N);
    q.lagP = lag;
    q.lagU = 2 * lag;
    // ring: planes in flight between the first producer and the last consumer, plus slack so that recycling a slot
    // never waits either
    q.ring = std::min(2 * lag + lag + 4, N);
    return q;
}

}  // namespace

// scratch of the fused kernel (doubles): 12 ring arrays (6 flux, drdT, drdS, 2 staged copies, 2 dissipation) +
// the work-queue head and the per-plane completion counters
size_t mega_ring_doubles(int N, int M, int nz, int eos, int energy) {
    const MegaPlan q = mega_plan(N, M, nz, eos, energy);
    return (size_t)12 * q.ring * M * nz;
}
size_t mega_sync_doubles(int N) { return ((size_t)4 * N + 2 + 1) / 2 + 1 + 16; }  // + 16 statistics words
size_t mega_stats_offset_doubles(int N) { return ((size_t)4 * N + 2 + 1) / 2 + 1; }

// ring + sync memory start at `scratch`; tables have been built (and `sync` zeroed) by launch_setup_tables
void launch_iso_mega(cudaStream_t s, const PreArgs& p0, const DiffArgs& d0, double* scratch, unsigned int* sync) {
    const int N = p0.g.N, M = p0.g.M, nz = p0.g.nz;
    const MegaPlan q = mega_plan(N, M, nz, p0.eos, d0.energy);
    MegaArgs m;
    m.p = p0;
    m.d = d0;
    m.ring = q.ring;
    m.cols = q.cols;
    m.pitch = q.pitch;
    m.nE = q.nE;
    m.nC = q.nC;
    m.nP = q.nP;
    m.nU = q.nU;
    m.sync = sync;
    m.stats = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(sync) + mega_stats_offset_doubles(N));
    m.lagP = q.lagP;
    m.lagU = q.lagU;
    m.fac_diss = 0.5 * d0.grav / d0.rho_0;  // diffusion.py (core) :19-21, Python float arithmetic
    m.gr = -d0.grav / d0.rho_0;             // isoneutral/diffusion.py:259,268
    const size_t rp = (size_t)q.ring * M * nz;
    double* w = scratch;
    for (int t = 0; t < 2; ++t)
        for (int k = 0; k < 3; ++k) {
            m.p.flux[t][k] = w;
            w += rp;
        }
    for (int t = 0; t < 2; ++t) {
        m.f.fe[t] = m.p.flux[t][0];
        m.f.fn[t] = m.p.flux[t][1];
        m.f.ft[t] = m.p.flux[t][2];
    }
    m.p.drdT = w;
    m.p.drdS = w + rp;
    m.p.stage[0] = w + 2 * rp;
    m.p.stage[1] = w + 3 * rp;
    m.f.diss[0] = w + 4 * rp;
    m.f.diss[1] = w + 5 * rp;
    m.p.with_flux = 1;
    m.p.with_stage = d0.energy;
    m.d.stage_x[0] = d0.energy ? m.p.stage[0] : nullptr;
    m.d.stage_x[1] = d0.energy ? m.p.stage[1] : nullptr;
    m.d.ntr = 2;
    m.d.skew = 0;
    m.p.two_rd = 2.0 / m.p.iso_dslope;
    m.p.m2c0 = -2.0 * m.p.iso_slopec / m.p.iso_dslope;
    m.p.s_max = (345.0 + m.p.iso_slopec / m.p.iso_dslope) * m.p.iso_dslope;

    const int total = (q.nE + q.nC + q.nP + q.nU) * (N + q.lagU);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define VB_MEGA(EOS_, EN_)                                                                                         \
    {                                                                                                              \
        auto kern = iso_mega_kernel<EOS_, EN_>;                                                                    \
        allow_big_smem(kern, 200 * 1024);                                                                          \
        int per_sm = 0;                                                                                            \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMegaBlock, q.smem) != cudaSuccess ||      \
            per_sm < 1)                                                                                            \
            per_sm = 1;                                                                                            \
        const unsigned grid = (unsigned)std::max(1, std::min(total, per_sm * sms));                                \
        kern<<<grid, kMegaBlock, q.smem, s>>>(m);                                                                  \
    }
#define VB_MEGA_EOS(EOS_)                      \
    if (d0.energy) VB_MEGA(EOS_, true) else VB_MEGA(EOS_, false)
    switch (p0.eos) {
    case 1: VB_MEGA_EOS(1) break;
    case 2: VB_MEGA_EOS(2) break;
    case 3: VB_MEGA_EOS(3) break;
    case 4: VB_MEGA_EOS(4) break;
    default: VB_MEGA_EOS(5) break;
    }
#undef VB_MEGA_EOS
#undef VB_MEGA
    count_launch();
    check_launch("iso_mega_kernel");
}

}  // namespace vb
