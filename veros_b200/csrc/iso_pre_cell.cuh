// The per-cell body of the slope / mixing-tensor / flux computation, shared by the stand-alone slope kernels
// (iso_pre.cu) and the fused persistent step kernel (iso_mega.cu).  See iso_pre.cu for the description of what one
// cell computes and why the arithmetic looks the way it does.
//
// Scratch arrays (TEOS-10 derivatives drdT/drdS, the six flux arrays) are addressed through `sc` / `sce`, the
// scratch index of cell (i, p) and of its east neighbour (i+1, p): the stand-alone kernels use the cell index itself
// (full-size scratch), the fused kernel a ring of a few x-planes that stays resident in L2.
#pragma once

#include "common.cuh"
#include "eos.cuh"
#include "flux_device.cuh"
#include "tables.cuh"

namespace vb {
namespace precell {

using strict::Divisor;
using strict::make_divisor;


using strict::Divisor;
using strict::make_divisor;

constexpr double kEps = 1e-20;  // isoneutral.py:28

// 1/x for normal finite x with 1/x normal: rcp.approx (>= 20 good bits) + one cubic Newton step.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// exp(u) for u in [-700, 700], relative error < 3e-16, no branches, no table:
//   u = n ln2 + r, |r| <= ln2/2;  exp(r) by its degree-12 Taylor polynomial (|r|^13/13! < 2e-16).
// (A 64-entry 2^(j/64) table with a degree-5 polynomial needs 6 fewer FP64 instructions but puts a
// dependent L1 load into each of the 16 taper chains of a cell; this kernel is latency bound.)
__device__ __forceinline__ double exp_fast(double u) {
    constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
    double t = fma(u, 0x1.71547652b82fep+0, kMagic);
    const int n = __double2loint(t);
    t -= kMagic;
    double r = fma(t, -0x1.62e42fefa3800p-1, u);
    r = fma(t, -0x1.ef35793c76730p-45, r);
    // Horner on purpose: an Estrin split (depth 4 instead of 12, 3 more instructions, 9 more live values)
    // measured 7-15 % slower -- this kernel pays for instructions and registers, not for chain depth.
    double p = 1.0 / 479001600.0;
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

struct Taper {
    double two_rd;   // 2 / iso_dslope
    double m2c0;     // -2 iso_slopec / iso_dslope
    double s_max;    // |s| beyond which exp(-2x) would overflow; the taper is exactly 0 there anyway
    // dm_taper (isoneutral.py:10-15): 0.5*(1+tanh(x)), x = (slopec-|s|)/dslope, as 1/(1+exp(-2x)).
    // "2q - 1" is rounded like a tanh value, so 1 + tanh(x) quantises to multiples of 2^-53 near -1
    // exactly as the reference's does (the taper is exactly 0 for x < -18.4).
    __device__ __forceinline__ double operator()(double s) const {
        const double sa = fmin(fabs(s), s_max);
        const double u = fma(sa, two_rd, m2c0);
        const double e = exp_fast(u);
        const double q = rcp_fast(1.0 + e);
        const double th = fma(2.0, q, -1.0);
        return fma(0.5, th, 0.5);
    }
};

// min(0, x) - eps with two FP64 instructions: x - |x| is 2x or 0 exactly.
__device__ __forceinline__ double neg_part_minus_eps(double x) { return fma(x - fabs(x), 0.5, -kEps); }

__device__ __forceinline__ void store_pair(double* base, double v0, double v1) {
    *reinterpret_cast<double2*>(base) = make_double2(v0, v1);
}

__device__ __forceinline__ double sel(bool m, double v) { return m ? v : 0.0; }


// One T cell (i, p = j*nz + k) of plane i.  T, S already point at time level tau.
// RING: the scratch arrays are written by other CTAs of the same launch (fused kernel): read them from L2, not
// through the non-coherent L1 path.
template <int EOS, bool FLUX, int FACES, bool RING>
__device__ __forceinline__ void pre_cell(const PreArgs& a, const Tables& tb, const double* __restrict__ T,
                                         const double* __restrict__ S, const int i, const int p, const size_t sc,
                                         const size_t sce) {
    constexpr bool doE = (FACES & 1) != 0, doN = (FACES & 2) != 0, doT = (FACES & 4) != 0;
    const int N = a.g.N, M = a.g.M, nz = a.g.nz;
    auto lds = [](const double* ptr) { return RING ? __ldcg(ptr) : __ldg(ptr); };
    const int j = p / nz;
    const int k = p - j * nz;
    const size_t plane = (size_t)M * nz;
    const size_t c = (size_t)i * plane + p;

    if (doT && k == nz - 1) a.K_33[c] = 0.0;  // isoneutral.py:225, whole array including ghost cells

    const bool inE = doE && (i >= 1 && i < N - 2 && j >= 2 && j < M - 2);
    const bool inN = doN && (i >= 2 && i < N - 2 && j >= 1 && j < M - 2);
    const bool inT = doT && (i >= 2 && i < N - 2 && j >= 2 && j < M - 2 && k < nz - 1);
    double fl[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};  // [tracer][east, north, top]

    // Masked faces.  Where maskU (east face), maskV (north face) or maskW (top face) of the cell is 0 the reference's
    // expressions reduce to exact zeros for finite fields: the masked gradient makes every slope of the face 0, so
    // Ai = taper * 0 = 0, the diffusivity sums carry the mask as a factor, and the fluxes multiply zeros
    // (isoneutral.py:64-95,115-131,196-222).  Such faces only STORE those zeros; a warp whose 32 cells need nothing
    // else skips all loads and arithmetic -- below the sea floor and on land that is every warp (about half of the
    // cells of a global grid with realistic bathymetry).
    const bool skip = !a.no_mask_skip;
    const bool needE = inE && !(skip && a.maskU[c] == 0);
    const bool needN = inN && !(skip && a.maskV[c] == 0);
    const bool needT = inT && !(skip && a.maskW[c] == 0);
    auto zero_face = [&](double* Ai, double* Kxx, bool all_kr) {
        double* out = Ai + c * 4;
        if (all_kr) {
            store_pair(out, 0.0, 0.0);
            store_pair(out + 2, 0.0, 0.0);
        } else {  // k = 0 of an east / north face: the kr = 0 entries are never written
            out[1] = 0.0;
            out[3] = 0.0;
        }
        Kxx[c] = 0.0;
    };
    if (inE && !needE) zero_face(a.Ai_ez, a.K_11, k >= 1);
    if (inN && !needN) zero_face(a.Ai_nz, a.K_22, k >= 1);
    if (inT && !needT) {
        zero_face(a.Ai_bx, a.K_33, true);
        store_pair(a.Ai_by + c * 4, 0.0, 0.0);
        store_pair(a.Ai_by + c * 4 + 2, 0.0, 0.0);
    }

    if (needE || needN || needT) {
        auto ld = [](const double* f, size_t cell) { return __ldg(f + cell * 3); };
        const bool hasKm = k >= 1, hasKp = k < nz - 1;
        const int km = hasKm ? -1 : 0, kp = hasKp ? 1 : 0;  // clamped neighbours (pad_z_edges)
        const size_t ce = c + plane, cn = c + nz, cw = c - plane, cs = c - nz;
        // At k = 0 the kr = 0 entries of Ai_ez / Ai_nz keep their old values and the fluxes read them
        // (times a zero difference).  They are cold DRAM lines: fetch them before anything else, not
        // next to their use behind the stores to the same array, where nothing could hide the miss.
        double Aez_old[2] = {0.0, 0.0}, Anz_old[2] = {0.0, 0.0};
        if (FLUX && !hasKm) {
            if (needE) { Aez_old[0] = a.Ai_ez[c * 4]; Aez_old[1] = a.Ai_ez[c * 4 + 2]; }
            if (needN) { Anz_old[0] = a.Ai_nz[c * 4]; Anz_old[1] = a.Ai_nz[c * 4 + 2]; }
        }
        // metric table entries of this level / row / plane (read-only path, L1 resident)
        struct { Divisor d4zt; double rdzw, dzw, pabs; } L1, L0;
        {
            const LevTab* l1 = tb.lev + k;
            const LevTab* l0 = tb.lev + k + km;
            L1.d4zt = ld_div(&l1->d4zt);
            const Divisor z1 = ld_div(&l1->dzw), z0 = ld_div(&l0->dzw);
            L1.rdzw = z1.ry; L1.dzw = z1.y; L1.pabs = __ldg(&l1->pabs);
            L0.rdzw = z0.ry; L0.dzw = z0.y; L0.pabs = 0.0; L0.d4zt = L1.d4zt;
        }
        struct { Divisor cdxu, dyu, cost, d4ytc; double cosu, facty; } Rj;
        {
            const RowTab* r = tb.row + j;
            Rj.dyu = ld_div(&r->dyu);
            Rj.cost = ld_div(&r->cost);
            Rj.d4ytc = ld_div(&r->d4ytc);
            const double2 cf = __ldg(reinterpret_cast<const double2*>(&r->cosu));
            Rj.cosu = cf.x; Rj.facty = cf.y;
            Rj.cdxu = ld_div(&tb.cell[(size_t)i * M + j].cdxu);
        }
        const Divisor d4xt = ld_div(&tb.xt[i].d4xt);
        const Taper taper = {a.two_rd, a.m2c0, a.s_max};

        // ---- this column: values, vertical differences and gradients ------------------------------
        const double Tc = ld(T, c), Sc = ld(S, c);
        const double Tm = ld(T, c + km), Sm = ld(S, c + km);
        const double Tp = ld(T, c + kp), Sp = ld(S, c + kp);
        const double dT0c = Tc - Tm, dS0c = Sc - Sm;  // level k-1 (0 at the bottom)
        const double dT1c = Tp - Tc, dS1c = Sp - Sc;  // level k   (0 at the surface)
        const bool mWc1 = hasKp && a.maskW[c], mWc0 = hasKm && a.maskW[c + km];
        const double w1c = sel(mWc1, L1.rdzw), w0c = sel(mWc0, L0.rdzw);
        const double gTz1c = dT1c * w1c, gSz1c = dS1c * w1c;  // dTdz(i,j,k)
        const double gTz0c = dT0c * w0c, gSz0c = dS0c * w0c;  // dTdz(i,j,k-1)

        double drTc, drSc;
        if (Eos<EOS>::kExpensive) {
            drTc = lds(a.drdT + sc);
            drSc = lds(a.drdS + sc);
        } else {
            eos_drho<EOS>(Sc, Tc, L1.pabs, drTc, drSc);
            const bool m = a.maskT[c];
            drTc = sel(m, drTc);
            drSc = sel(m, drSc);
        }
        const double Kc = __ldg(a.K_iso + c);
        const double Kcm = __ldg(a.K_iso + c + km);
        const double rdzt4 = L1.d4zt.ry;

        // The triad (density point (i,j,k), east face of the cell, level k below the top face) enters both the east
        // face (ip = 0, kr = 1) and the top face (ip = 1, kr = 0) with bit-identical slope and taper; likewise in y.
        // When one launch computes all faces the top face reuses them: 14 tapers per cell instead of 16.
        double sh_sx = 0.0, sh_tsx = 0.0, sh_sy = 0.0, sh_tsy = 0.0;
        // ---- east face: Ai_ez, K_11 (isoneutral.py:100-132) and flux_east (diffusion.py:25-47) ------
        double Te = 0.0, Se = 0.0, Tpe = 0.0, Spe = 0.0;  // (i+1,j,k), (i+1,j,k+1)
        if (needE || needT) {
            Te = ld(T, ce);
            Se = ld(S, ce);
            Tpe = ld(T, ce + kp);
            Spe = ld(S, ce + kp);
        }
        const double dTxc = Te - Tc, dSxc = Se - Sc;  // raw east differences at level k
        const double mrdx = sel(a.maskU[c] != 0, Rj.cdxu.ry);
        const double gTxc = dTxc * mrdx, gSxc = dSxc * mrdx;  // dTdx(i,j,k)
        if (needE) {
            const double Tme = ld(T, ce + km), Sme = ld(S, ce + km);
            const double dT0e = Te - Tme, dS0e = Se - Sme, dT1e = Tpe - Te, dS1e = Spe - Se;
            const bool mWe1 = hasKp && a.maskW[ce], mWe0 = hasKm && a.maskW[ce + km];
            const double w1e = sel(mWe1, L1.rdzw), w0e = sel(mWe0, L0.rdzw);
            double drTe, drSe;
            if (Eos<EOS>::kExpensive) {
                drTe = lds(a.drdT + sce);
                drSe = lds(a.drdS + sce);
            } else {
                eos_drho<EOS>(Se, Te, L1.pabs, drTe, drSe);
                const bool m = a.maskT[ce];
                drTe = sel(m, drTe);
                drSe = sel(m, drSe);
            }
            const double Ke = __ldg(a.K_iso + ce), Kem = __ldg(a.K_iso + ce + km);
            const double diffloc = hasKm ? 0.25 * (((Kc + Kcm) + Ke) + Kem) : 0.5 * (Kc + Ke);
            const bool mU = a.maskU[c] != 0;
            const double wz[2] = {sel(mU && hasKm, L0.dzw), sel(mU, L1.dzw)};  // dzw[k+kr-1] * maskU
            const double gTz[2][2] = {{gTz0c, gTz1c}, {dT0e * w0e, dT1e * w1e}};  // [ip][kr]
            const double gSz[2][2] = {{gSz0c, gSz1c}, {dS0e * w0e, dS1e * w1e}};
            double A[2][2];
            double sumz = 0.0;
#pragma unroll
            for (int kr = 0; kr < 2; ++kr) {
#pragma unroll
                for (int ip = 0; ip < 2; ++ip) {
                    const double dT_ = ip ? drTe : drTc, dS_ = ip ? drSe : drSc;
                    const double drodxe = fma(dS_, gSxc, dT_ * gTxc);
                    const double drodze = fma(dS_, gSz[ip][kr], dT_ * gTz[ip][kr]);
                    const double sxe = -drodxe * rcp_fast(neg_part_minus_eps(drodze));
                    const double tp = taper(sxe);
                    sumz = fma(wz[kr], fmax(a.K_iso_steep, diffloc * tp), sumz);
                    A[ip][kr] = tp * sxe;  // maskU is already in gTxc/gSxc
                    if (kr == 1 && ip == 0) { sh_sx = sxe; sh_tsx = A[ip][kr]; }
                }
            }
            double* out = a.Ai_ez + c * 4;
            if (hasKm) {
                store_pair(out, A[0][0], A[0][1]);
                store_pair(out + 2, A[1][0], A[1][1]);
            } else {  // k = 0: the kr = 0 entries are never written (isoneutral.py:113-131, ki = 1)
                out[1] = A[0][1];
                out[3] = A[1][1];
            }
            const double K11 = sumz * rdzt4;
            a.K_11[c] = K11;
            if (FLUX) {
                // at k = 0 the kr = 0 entries of Ai_ez keep their old values; they multiply a zero difference
                const double A00 = hasKm ? A[0][0] : Aez_old[0], A10 = hasKm ? A[1][0] : Aez_old[1];
                fl[0][0] = flux_face(diffloc, A00, A[0][1], A10, A[1][1], dT0c, dT1c, dT0e, dT1e, dTxc, L1.d4zt, Rj.cdxu, K11);
                fl[1][0] = flux_face(diffloc, A00, A[0][1], A10, A[1][1], dS0c, dS1c, dS0e, dS1e, dSxc, L1.d4zt, Rj.cdxu, K11);
            }
        }

        // ---- north face: Ai_nz, K_22 (isoneutral.py:137-168) and flux_north (diffusion.py:52-77) ----
        double Tn = 0.0, Sn = 0.0, Tpn = 0.0, Spn = 0.0;  // (i,j+1,k), (i,j+1,k+1)
        if (needN || needT) {
            Tn = ld(T, cn);
            Sn = ld(S, cn);
            Tpn = ld(T, cn + kp);
            Spn = ld(S, cn + kp);
        }
        const double dTyc = Tn - Tc, dSyc = Sn - Sc;
        const double mrdy = sel(a.maskV[c] != 0, Rj.dyu.ry);
        const double gTyc = dTyc * mrdy, gSyc = dSyc * mrdy;  // dTdy(i,j,k)
        if (needN) {
            const double Tmn = ld(T, cn + km), Smn = ld(S, cn + km);
            const double dT0n = Tn - Tmn, dS0n = Sn - Smn, dT1n = Tpn - Tn, dS1n = Spn - Sn;
            const bool mWn1 = hasKp && a.maskW[cn], mWn0 = hasKm && a.maskW[cn + km];
            const double w1n = sel(mWn1, L1.rdzw), w0n = sel(mWn0, L0.rdzw);
            double drTn, drSn;
            if (Eos<EOS>::kExpensive) {
                drTn = lds(a.drdT + sc + nz);
                drSn = lds(a.drdS + sc + nz);
            } else {
                eos_drho<EOS>(Sn, Tn, L1.pabs, drTn, drSn);
                const bool m = a.maskT[cn];
                drTn = sel(m, drTn);
                drSn = sel(m, drSn);
            }
            const double Kn = __ldg(a.K_iso + cn), Knm = __ldg(a.K_iso + cn + km);
            const double diffloc = hasKm ? 0.25 * (((Kc + Kcm) + Kn) + Knm) : 0.5 * (Kc + Kn);
            const bool mV = a.maskV[c] != 0;
            const double wz[2] = {sel(mV && hasKm, L0.dzw), sel(mV, L1.dzw)};
            const double gTz[2][2] = {{gTz0c, gTz1c}, {dT0n * w0n, dT1n * w1n}};  // [jp][kr]
            const double gSz[2][2] = {{gSz0c, gSz1c}, {dS0n * w0n, dS1n * w1n}};
            double A[2][2];
            double sumz = 0.0;
#pragma unroll
            for (int kr = 0; kr < 2; ++kr) {
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    const double dT_ = jp ? drTn : drTc, dS_ = jp ? drSn : drSc;
                    const double drodyn = fma(dS_, gSyc, dT_ * gTyc);
                    const double drodzn = fma(dS_, gSz[jp][kr], dT_ * gTz[jp][kr]);
                    const double syn = -drodyn * rcp_fast(neg_part_minus_eps(drodzn));
                    const double tp = taper(syn);
                    sumz = fma(wz[kr], fmax(a.K_iso_steep, diffloc * tp), sumz);
                    A[jp][kr] = tp * syn;
                    if (kr == 1 && jp == 0) { sh_sy = syn; sh_tsy = A[jp][kr]; }
                }
            }
            double* out = a.Ai_nz + c * 4;
            if (hasKm) {
                store_pair(out, A[0][0], A[0][1]);
                store_pair(out + 2, A[1][0], A[1][1]);
            } else {
                out[1] = A[0][1];
                out[3] = A[1][1];
            }
            const double K22 = sumz * rdzt4;
            a.K_22[c] = K22;
            if (FLUX) {
                const double A00 = hasKm ? A[0][0] : Anz_old[0], A10 = hasKm ? A[1][0] : Anz_old[1];
                fl[0][1] = strict::mul(Rj.cosu, flux_face(diffloc, A00, A[0][1], A10, A[1][1], dT0c, dT1c, dT0n, dT1n,
                                                          dTyc, L1.d4zt, Rj.dyu, K22));
                fl[1][1] = strict::mul(Rj.cosu, flux_face(diffloc, A00, A[0][1], A10, A[1][1], dS0c, dS1c, dS0n, dS1n,
                                                          dSyc, L1.d4zt, Rj.dyu, K22));
            }
        }

        // ---- top face: Ai_bx, Ai_by, K_33 (isoneutral.py:173-225) and flux_top (diffusion.py:85-111) --
        if (needT) {
            struct { Divisor dyu; double cosu, facty; } Rs;
            {
                const RowTab* r = tb.row + j - 1;
                Rs.dyu = ld_div(&r->dyu);
                const double2 cf = __ldg(reinterpret_cast<const double2*>(&r->cosu));
                Rs.cosu = cf.x; Rs.facty = cf.y;
            }
            const double r_cdxu_w = __ldg(&tb.cell[(size_t)(i - 1) * M + j].cdxu.ry);
            const double Tw = ld(T, cw), Sw = ld(S, cw), Tpw = ld(T, cw + 1), Spw = ld(S, cw + 1);
            const double Ts = ld(T, cs), Ss = ld(S, cs), Tps = ld(T, cs + 1), Sps = ld(S, cs + 1);
            // raw differences [ip|jp][kr]: tr(i+ip,j,k+kr) - tr(i-1+ip,j,k+kr) and the same in y
            const double dTx[2][2] = {{Tc - Tw, Tp - Tpw}, {dTxc, Tpe - Tp}};
            const double dSx[2][2] = {{Sc - Sw, Sp - Spw}, {dSxc, Spe - Sp}};
            const double dTy[2][2] = {{Tc - Ts, Tp - Tps}, {dTyc, Tpn - Tp}};
            const double dSy[2][2] = {{Sc - Ss, Sp - Sps}, {dSyc, Spn - Sp}};
            // metric factors with the U/V masks folded in
            const double mx[2][2] = {{sel(a.maskU[cw] != 0, r_cdxu_w), sel(a.maskU[cw + 1] != 0, r_cdxu_w)},
                                     {mrdx, sel(a.maskU[c + 1] != 0, Rj.cdxu.ry)}};
            const double my[2][2] = {{sel(a.maskV[cs] != 0, Rs.dyu.ry), sel(a.maskV[cs + 1] != 0, Rs.dyu.ry)},
                                     {mrdy, sel(a.maskV[c + 1] != 0, Rj.dyu.ry)}};
            double drTu, drSu;  // (i,j,k+1)
            if (Eos<EOS>::kExpensive) {
                drTu = lds(a.drdT + sc + 1);
                drSu = lds(a.drdS + sc + 1);
            } else {
                eos_drho<EOS>(Sp, Tp, __ldg(&tb.lev[k + 1].pabs), drTu, drSu);
                const bool m = a.maskT[c + 1];
                drTu = sel(m, drTu);
                drSu = sel(m, drSu);
            }
            const double KcW = sel(mWc1, Kc);                                 // K_iso * maskW
            const double cx[2] = {__ldg(&tb.xt[i - 1].dxu) * KcW, __ldg(&tb.xt[i].dxu) * KcW};
            const double cy[2] = {Rs.facty * KcW, Rj.facty * KcW};
            double Ax[2][2], Ay[2][2];
            double sumx = 0.0, sumy = 0.0;
#pragma unroll
            for (int kr = 0; kr < 2; ++kr) {
                const double dT_ = kr ? drTu : drTc, dS_ = kr ? drSu : drSc;
                const double drodzb = fma(dS_, gSz1c, dT_ * gTz1c);
                const double nrden = -rcp_fast(neg_part_minus_eps(drodzb));
#pragma unroll
                for (int ip = 0; ip < 2; ++ip) {
                    double sxb, ts;
                    if (doE && kr == 0 && ip == 1) {  // = the east face's (ip = 0, kr = 1) triad
                        sxb = sh_sx;
                        ts = sh_tsx;
                    } else {
                        const double drodxb = fma(dS_, dSx[ip][kr] * mx[ip][kr], dT_ * (dTx[ip][kr] * mx[ip][kr]));
                        sxb = drodxb * nrden;
                        ts = taper(sxb) * sxb;
                    }
                    sumx = fma(cx[ip], ts * sxb, sumx);
                    Ax[ip][kr] = sel(mWc1, ts);
                }
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    double syb, ts;
                    if (doN && kr == 0 && jp == 1) {  // = the north face's (jp = 0, kr = 1) triad
                        syb = sh_sy;
                        ts = sh_tsy;
                    } else {
                        const double drodyb = fma(dS_, dSy[jp][kr] * my[jp][kr], dT_ * (dTy[jp][kr] * my[jp][kr]));
                        syb = drodyb * nrden;
                        ts = taper(syb) * syb;
                    }
                    sumy = fma(cy[jp], ts * syb, sumy);
                    Ay[jp][kr] = sel(mWc1, ts);
                }
            }
            store_pair(a.Ai_bx + c * 4, Ax[0][0], Ax[0][1]);
            store_pair(a.Ai_bx + c * 4 + 2, Ax[1][0], Ax[1][1]);
            store_pair(a.Ai_by + c * 4, Ay[0][0], Ay[0][1]);
            store_pair(a.Ai_by + c * 4 + 2, Ay[1][0], Ay[1][1]);
            a.K_33[c] = fma(sumx, d4xt.ry, sumy * Rj.d4ytc.ry);
            if (FLUX) {
                fl[0][2] = flux_top(Kc, Ax, Ay, dTx, dTy, Rs.cosu, Rj.cosu, Rj.cost, d4xt, Rj.d4ytc);
                fl[1][2] = flux_top(Kc, Ax, Ay, dSx, dSy, Rs.cosu, Rj.cosu, Rj.cost, d4xt, Rj.d4ytc);
            }
        }
    }
    if (FLUX) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (doE) a.flux[t][0][sc] = fl[t][0];
            if (doN) a.flux[t][1][sc] = fl[t][1];
            if (doT) a.flux[t][2][sc] = fl[t][2];
        }
    }
}

}  // namespace precell
}  // namespace vb
