// The phases of the tracer update on one tile of whole columns (explicit flux divergence, implicit K_33 column solve,
// tendencies, dissipation), shared by update_kernel (iso_diffusion.cu) and the fused persistent step kernel
// (iso_mega.cu).  See iso_diffusion.cu for what each phase computes.
//
// Scratch addressing: the flux arrays, the staged int_drhodX copies and the T-point dissipation live in scratch.
// TileGeom carries the scratch offsets of the tile in plane i (sb_c), i-1 (sb_w) and i+1 (sb_e) and of its
// dissipation slot (db): the stand-alone kernel uses full-size scratch (offset = cell index), the fused kernel a ring
// of a few x-planes and a per-CTA dissipation slot.  COHERENT: scratch and K_33 were written by other CTAs of the
// same launch -- read them from L2 (ld.global.cg), not through the non-coherent L1 path.
#pragma once

#include "common.cuh"
#include "strict.cuh"
#include "tdma_device.cuh"

namespace vb {
namespace upd {

using strict::add;
using strict::Divisor;
using strict::make_divisor;
using strict::mul;
using strict::sub;

template <bool COHERENT>
__device__ __forceinline__ double lds(const double* p) { return COHERENT ? __ldcg(p) : __ldg(p); }

struct Scratch {
    const double *fe[2], *fn[2], *ft[2];
    double* diss[2];  // T-point dissipation of each tracer (ENERGY only)
};

constexpr int kUpdBlock = 128;

// ---- pieces of the update shared by the plain and the pipelined kernel ---------------------------------
struct LevelTabs {        // per level, once per CTA
    Divisor* ddzt;        // dzt[k]
    Divisor* ddzw;        // dzw[k]
    double* dt_dzw;       // dt_tracer / dzw[k]
};
struct TileBuf {          // one tile of whole columns in shared memory
    double *L, *D, *U, *R[2];
    Divisor *dcdxt, *dcdyt;  // cost[j]*dxt[i], cost[j]*dyt[j] per column
    int* ksv;                // kbot - 1 per column
};
struct TileGeom {
    int i, j0, ncols, ncells;
    size_t base;              // cell index of the tile's first cell
    size_t sb_c, sb_w, sb_e;  // scratch offsets of the tile's first cell in planes i, i-1, i+1
    size_t db;                // offset of the tile's first cell in the dissipation scratch
    bool i_int;
};
struct UpdConst {
    int N, M, nz, pitch, tau, taup1;
    size_t plane;
    double dt, fac_diss, gr;
    Divisor ddt;
};

__device__ __forceinline__ TileBuf tile_buf_at(double* p, int cols, int pitch, int ntr) {
    TileBuf b;
    const int tile = cols * pitch;
    b.L = p;
    b.D = b.L + tile;
    b.U = b.D + tile;
    b.R[0] = b.U + tile;
    b.R[1] = b.R[0] + (ntr > 1 ? tile : 0);
    b.dcdxt = reinterpret_cast<Divisor*>(b.R[0] + (size_t)ntr * tile);
    b.dcdyt = b.dcdxt + cols;
    b.ksv = reinterpret_cast<int*>(b.dcdyt + cols);
    return b;
}

__device__ __forceinline__ TileGeom tile_geom(const UpdConst& u, int i, int j0, int cols) {
    TileGeom g;
    g.i = i;
    g.j0 = j0;
    g.ncols = min(cols, (u.M - 1) - j0);
    g.ncells = g.ncols * u.nz;
    g.base = (size_t)i * u.plane + (size_t)j0 * u.nz;
    g.sb_c = g.base;  // full-size scratch; the fused kernel overrides these with ring offsets
    g.sb_w = g.base - u.plane;
    g.sb_e = g.base + u.plane;
    g.db = g.base;
    g.i_int = (i >= 2 && i < u.N - 2);
    return g;
}

__device__ __forceinline__ void fill_level_tabs(const DiffArgs& a, const LevelTabs& lv, int nz, double dt, int tid, int nthr) {
    for (int k = tid; k < nz; k += nthr) {
        lv.ddzt[k] = make_divisor(a.g.dzt[k]);
        lv.ddzw[k] = make_divisor(a.g.dzw[k]);
        lv.dt_dzw[k] = strict::div(dt, a.g.dzw[k]);
    }
}

__device__ __forceinline__ void fill_tile_tabs(const DiffArgs& a, const UpdConst& u, const TileBuf& b, const TileGeom& g,
                                               int tid, int nthr) {
    for (int q = tid; q < g.ncols; q += nthr) {
        const int j = g.j0 + q;
        b.dcdxt[q] = make_divisor(mul(a.g.cost[j], a.g.dxt[g.i]));
        b.dcdyt[q] = make_divisor(mul(a.g.cost[j], a.g.dyt[j]));
        b.ksv[q] = a.kbot[g.i * u.M + j] - 1;
    }
}

// phase B: explicit flux divergence, tracer + tendency update, right-hand sides, matrix, T-point dissipation
template <int NTR, bool SKEW, bool ENERGY, bool COHERENT>
__device__ __forceinline__ void phase_b(const DiffArgs& a, const Scratch& f, const UpdConst& u, const LevelTabs& lv,
                                        const TileBuf& b, const TileGeom& g, int tid, int nthr) {
    const int N = u.N, M = u.M, nz = u.nz, pitch = u.pitch, tau = u.tau, taup1 = u.taup1;
    const size_t plane = u.plane;
    const double dt = u.dt, fac_diss = u.fac_diss;
    (void)N;
    // ---- phase B --------------------------------------------------------------------------------
    // Every global load of a (cell, tracer) pair is issued before the first dependent store: the
    // compiler must keep loads behind earlier stores that might alias, so interleaving them would
    // serialise four memory round trips per cell.
    for (int idx = tid; idx < g.ncells; idx += nthr) {
        const int q = idx / nz, k = idx - q * nz;
        const int j = g.j0 + q;
        const bool interior = g.i_int && j >= 2 && j < M - 2;
        if (!interior && !ENERGY) continue;
        const size_t c = g.base + idx;
        const int s = q * pitch + k;
        // Dry cells (fused step, masks consistent with kbot as veros/core/numerics.py:200-221 builds them): the explicit
        // tendency is maskT * (...) = 0, so tracer and tendency keep their values bit for bit; every flux around the
        // cell is masked, so its dissipation is 0; the column solve starts above it.  Nothing to load, nothing to
        // store but that zero (phase D of the wet cell above may read it).
        if (a.dry_skip && a.maskT[c] == 0) {
            if (ENERGY) {
#pragma unroll
                for (int t = 0; t < NTR; ++t) f.diss[t][g.db + idx] = 0.0;
            }
            continue;
        }
        const double mT = interior ? (double)a.maskT[c] : 0.0;
        double k33 = 0.0, k33m = 0.0;
        if (!SKEW && interior) {
            k33 = (k < nz - 1) ? lds<COHERENT>(a.K_33 + c) : 0.0;
            k33m = (k > 0) ? lds<COHERENT>(a.K_33 + c - 1) : 0.0;
        }
#pragma unroll
        for (int t = 0; t < NTR; ++t) {
            // loads
            const size_t sc = g.sb_c + idx, sw = g.sb_w + idx, se = g.sb_e + idx;
            const double fe_c = lds<COHERENT>(f.fe[t] + sc), fe_w = lds<COHERENT>(f.fe[t] + sw);
            const double fn_c = lds<COHERENT>(f.fn[t] + sc), fn_s = lds<COHERENT>(f.fn[t] + sc - nz);
            double ft_c = 0.0, ft_m = 0.0, dtr_old = 0.0, tr_old = 0.0;
            if (interior) {
                ft_c = lds<COHERENT>(f.ft[t] + sc);
                ft_m = k > 0 ? lds<COHERENT>(f.ft[t] + sc - 1) : 0.0;
                dtr_old = a.t[t].dtracer[c];
                tr_old = a.t[t].tr[c * 3 + taup1];
            }
            double xc = 0.0, xe = 0.0, xw = 0.0, xn = 0.0, xs = 0.0;
            if (ENERGY) {
                // int_drhodX[..., tau]: the step's contiguous copy if there is one, else the strided original
                if (a.stage_x[t] != nullptr) {  // the step's contiguous copy (scratch addressing)
                    const double* __restrict__ X = a.stage_x[t];
                    xc = lds<COHERENT>(X + sc);
                    xe = lds<COHERENT>(X + se);
                    xw = lds<COHERENT>(X + sw);
                    xn = lds<COHERENT>(X + sc + nz);
                    xs = lds<COHERENT>(X + sc - nz);
                } else {  // the strided original
                    const double* __restrict__ X = a.t[t].int_drhodX + tau;
                    xc = __ldg(X + c * 3);
                    xe = __ldg(X + (c + plane) * 3);
                    xw = __ldg(X + (c - plane) * 3);
                    xn = __ldg(X + (c + nz) * 3);
                    xs = __ldg(X + (c - nz) * 3);
                }
            }
            // arithmetic + stores
            if (interior) {
                double e = mul(mT, add(strict::div(sub(fe_c, fe_w), b.dcdxt[q]), strict::div(sub(fn_c, fn_s), b.dcdyt[q])));
                if (k == 0)
                    e = add(e, strict::div(mul(mT, ft_c), lv.ddzt[0]));
                else
                    e = add(e, strict::div(mul(mT, sub(ft_c, ft_m)), lv.ddzt[k]));
                a.t[t].dtracer[c] = add(dtr_old, e);        // diffusion.py:196
                const double v = add(tr_old, mul(dt, e));  // diffusion.py:197
                a.t[t].tr[c * 3 + taup1] = v;
                if (!SKEW) b.R[t][s] = v;
            }
            if (ENERGY) {  // compute_dissipation, veros/core/diffusion.py:15-35 (on [1:-1, 1:-1])
                const double gx = add(mul(sub(xe, xc), fe_c), mul(sub(xc, xw), fe_w));
                const double gy = add(mul(sub(xn, xc), fn_c), mul(sub(xc, xs), fn_s));
                f.diss[t][g.db + idx] = add(strict::div(mul(fac_diss, gx), b.dcdxt[q]), strict::div(mul(fac_diss, gy), b.dcdyt[q]));
            }
        }
        if (!SKEW && interior) {  // _calc_implicit_part, diffusion.py:149-164
            const int ks = b.ksv[q];
            const double del = (k < nz - 1) ? mul(lv.dt_dzw[k], k33) : 0.0;
            const double delm = (k > 0) ? mul(lv.dt_dzw[k - 1], k33m) : 0.0;
            double diag;
            if (k == ks)
                diag = add(1.0, strict::div(del, lv.ddzt[k]));  // b_tri_edge
            else if (k == nz - 1)
                diag = add(1.0, strict::div(delm, lv.ddzt[k]));
            else
                diag = add(1.0, strict::div(add(del, delm), lv.ddzt[k]));
            b.D[s] = diag;
            b.U[s] = (k < nz - 1) ? strict::div(-del, lv.ddzt[k]) : 0.0;
            if (k > 0) b.L[s - 1] = (k > ks) ? strict::div(-delm, lv.ddzt[k]) : 0.0;
        }
    }
}

// phase C: one thread per water column, dgtsv on all right-hand sides
template <int NTR>
__device__ __forceinline__ void phase_c(const UpdConst& u, const TileBuf& b, const TileGeom& g, int tid, int nthr) {
    if (!g.i_int) return;
    for (int q = tid; q < g.ncols; q += nthr) {
        const int j = g.j0 + q;
        const int ks = b.ksv[q];
        if (j >= 2 && j < u.M - 2 && ks >= 0) {
            const int o = q * u.pitch;
            dgtsv_column<NTR>(ks, u.nz, b.L + o, b.D + o, b.U + o, b.R[0] + o, b.R[NTR - 1] + o);
        }
    }
}

// phase D: implicit result, tendency, dissipation on the W grid
template <int NTR, bool SKEW, bool ENERGY, bool COHERENT>
__device__ __forceinline__ void phase_d(const DiffArgs& a, const Scratch& f, const UpdConst& u, const LevelTabs& lv,
                                        const TileBuf& b, const TileGeom& g, int tid, int nthr) {
    const int M = u.M, nz = u.nz, pitch = u.pitch, tau = u.tau, taup1 = u.taup1;
    const double gr = u.gr;
    const Divisor ddt = u.ddt;
    // ---- phase D ----------------------------------------------------------------------------------
    for (int idx = tid; idx < g.ncells; idx += nthr) {
        const int q = idx / nz, k = idx - q * nz;
        const int j = g.j0 + q;
        const size_t c = g.base + idx;
        const int s = q * pitch + k;
        const bool interior = g.i_int && j >= 2 && j < M - 2;
        const int ks = b.ksv[q];
        const bool land = ks >= 0;
        const bool up = k < nz - 1;
        const bool solved = !SKEW && interior && land && k >= ks;
        if (!ENERGY && !solved) continue;
        if (a.dry_skip && a.maskT[c] == 0) continue;  // dissipation_on_wgrid and the vertical term are exact zeros there
        // loads
        double P = 0.0, k33 = 0.0, mW = 0.0;
        double old[NTR], dtr_mid[NTR], d0[NTR], d1[NTR], x0[NTR], x1[NTR], ftc[NTR];
        if (ENERGY) {
            P = a.P_diss[c];
            if (interior && up) {
                k33 = lds<COHERENT>(a.K_33 + c);
                mW = (double)a.maskW[c];
            }
        }
#pragma unroll
        for (int t = 0; t < NTR; ++t) {
            old[t] = dtr_mid[t] = d0[t] = d1[t] = x0[t] = x1[t] = ftc[t] = 0.0;
            if (solved) {
                old[t] = a.t[t].tr[c * 3 + taup1];
                dtr_mid[t] = a.t[t].dtracer[c];
            }
            if (ENERGY) {
                d0[t] = f.diss[t][g.db + idx];
                if (up) d1[t] = f.diss[t][g.db + idx + 1];
                if (interior && up) {
                    const size_t sc = g.sb_c + idx;
                    if (a.stage_x[t] != nullptr) {
                        x0[t] = lds<COHERENT>(a.stage_x[t] + sc);
                        x1[t] = lds<COHERENT>(a.stage_x[t] + sc + 1);
                    } else {
                        const double* __restrict__ X = a.t[t].int_drhodX + tau;
                        x0[t] = __ldg(X + c * 3);
                        x1[t] = __ldg(X + (c + 1) * 3);
                    }
                    ftc[t] = lds<COHERENT>(f.ft[t] + sc);
                }
            }
        }
        // arithmetic + stores
#pragma unroll
        for (int t = 0; t < NTR; ++t) {
            if (solved) {  // where(water_mask, sol, tr); diffusion.py:168,203-204
                const double nw = b.R[t][s];
                a.t[t].dtracer[c] = add(dtr_mid[t], strict::div(sub(nw, old[t]), ddt));
                a.t[t].tr[c * 3 + taup1] = nw;
            }
            if (ENERGY) {
                // dissipation_on_wgrid, veros/core/diffusion.py:41-62
                double dw;
                if (up) {
                    const double m = mul(0.5, add(d0[t], d1[t]));
                    const double edge = (land && k == ks) ? 1.0 : 0.0, water = (land && k > ks) ? 1.0 : 0.0;
                    const double dzw_pad = lv.ddzw[k > 0 ? k - 1 : 0].y;
                    dw = add(mul(add(m, mul(0.5, strict::div(mul(d0[t], dzw_pad), lv.ddzw[k]))), edge), mul(m, water));
                } else {
                    dw = mul(d0[t], land ? 1.0 : 0.0);
                }
                P = add(P, dw);  // diffusion.py:246-249
                if (interior && up) {  // diffusion.py:254-279
                    const double fxa = strict::div(add(-x1[t], x0[t]), lv.ddzw[k]);
                    double v;
                    if (SKEW) {
                        v = mul(mul(mul(gr, fxa), ftc[t]), mW);
                    } else {
                        // tr[taup1] after the update: R holds it for every interior cell of the tile
                        const double dtr = sub(b.R[t][s + 1], b.R[t][s]);
                        v = mul(mul(gr, fxa), add(mul(ftc[t], mW), mul(strict::div(mul(k33, dtr), lv.ddzw[k]), mW)));
                    }
                    P = add(P, v);
                }
            }
        }
        if (ENERGY) a.P_diss[c] = P;
    }
}

__device__ __forceinline__ UpdConst upd_const(const DiffArgs& a, int pitch, double fac_diss, double gr) {
    UpdConst u;
    u.N = a.g.N;
    u.M = a.g.M;
    u.nz = a.g.nz;
    u.pitch = pitch;
    u.tau = *a.tau;
    u.taup1 = *a.taup1;
    u.plane = (size_t)a.g.M * a.g.nz;
    u.dt = a.dt_tracer;
    u.fac_diss = fac_diss;
    u.gr = gr;
    u.ddt = make_divisor(a.dt_tracer);
    return u;
}


}  // namespace upd
}  // namespace vb
