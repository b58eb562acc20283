// isoneutral_diffusion_pre (veros/core/isoneutral/isoneutral.py:18-229): density-triad slopes,
// Ai_ez / Ai_nz / Ai_bx / Ai_by and the mixing tensor K_11 / K_22 / K_33, one pass over the grid.
//
// One thread owns one T cell (i,j,k) and produces everything the reference stores at that index:
// the east-face, north-face and top-face triads (16 slopes).  Threads are laid out along the
// flattened (j,k) index of an x-plane, i.e. along memory, so every load and store of a warp is
// contiguous; the 2x2 blocks of the Ai_* arrays are written as two 16-byte stores per thread.
// Nothing intermediate (gradients, drdT/drdS, diffloc, sums) ever goes to HBM, except drdT/drdS
// for the 48-term TEOS-10 equation of state, which a small pre-pass evaluates once per cell.
//
// Arithmetic: this kernel is FP64-pipe bound, not HBM bound (16 divisions + 16 tanh per cell in
// the reference formulation).  Divisions by grid metrics are multiplications by reciprocals and
// 0.5*(1+tanh(x)) is evaluated through one exp and one reciprocal; results agree with the
// reference to ~1e-15 (tests/test_gpu_parity.py), they are not bit-identical (neither are NumPy's
// SIMD tanh and libm's).
#include "common.cuh"
#include "eos.cuh"

namespace vb {

namespace {

constexpr double kEps = 1e-20;  // isoneutral.py:28

struct TaperParams {
    double c0;       // iso_slopec / iso_dslope
    double rdslope;  // 1 / iso_dslope
};

// dm_taper (isoneutral.py:10-15): 0.5*(1+tanh(x)), x = (-|s| + slopec)/dslope.
// tanh(x) = 2/(1+exp(-2x)) - 1.  The intermediate "2q - 1" is rounded like the reference's tanh
// value so that 1 + tanh(x) quantises to multiples of 2^-53 near -1 exactly as NumPy's does
// (the taper is exactly 0 for x < -18.4).
__device__ __forceinline__ double dm_taper(double s, const TaperParams& tp) {
    const double x = fma(-fabs(s), tp.rdslope, tp.c0);
    const double e = exp(-2.0 * x);
    const double q = 1.0 / (1.0 + e);
    const double th = fma(2.0, q, -1.0);
    return 0.5 * (1.0 + th);
}

__device__ __forceinline__ void store_pair(double* base, double v0, double v1) {
    *reinterpret_cast<double2*>(base) = make_double2(v0, v1);
}

}  // namespace

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

template <int EOS>
__global__ void __launch_bounds__(128)
iso_pre_kernel(const PreArgs a) {
    const int N = a.g.N, M = a.g.M, nz = a.g.nz;
    const int i = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M * nz) return;
    const int j = p / nz;
    const int k = p - j * nz;
    const size_t plane = (size_t)M * nz;
    const size_t c = (size_t)i * plane + p;

    if (k == nz - 1) a.K_33[c] = 0.0;  // isoneutral.py:225, whole array including ghost cells

    const bool inE = (i >= 1 && i < N - 2 && j >= 2 && j < M - 2);
    const bool inN = (i >= 2 && i < N - 2 && j >= 1 && j < M - 2);
    const bool inT = (i >= 2 && i < N - 2 && j >= 2 && j < M - 2 && k < nz - 1);
    if (!(inE || inN || inT)) return;

    const int tau = *a.tau;
    const double* __restrict__ T = a.temp + tau;
    const double* __restrict__ S = a.salt + tau;
    auto ld = [](const double* f, size_t cell) { return __ldg(f + cell * 3); };

    const bool hasKm = k >= 1, hasKp = k < nz - 1;
    const size_t ce = c + plane, cn = c + nz, cw = c - plane, cs = c - nz;

    // ---- tracer values ------------------------------------------------------------------------
    const double Tc = ld(T, c), Sc = ld(S, c);
    const double Tkm = hasKm ? ld(T, c - 1) : 0.0, Skm = hasKm ? ld(S, c - 1) : 0.0;
    const double Tkp = hasKp ? ld(T, c + 1) : 0.0, Skp = hasKp ? ld(S, c + 1) : 0.0;

    const double rdz1 = hasKp ? 1.0 / __ldg(a.g.dzw + k) : 0.0;      // level k   (between k and k+1)
    const double rdz0 = hasKm ? 1.0 / __ldg(a.g.dzw + k - 1) : 0.0;  // level k-1
    const double dzw1 = __ldg(a.g.dzw + k);
    const double dzw0 = hasKm ? __ldg(a.g.dzw + k - 1) : 0.0;
    const double pk = fabs(__ldg(a.g.zt + k));

    // vertical gradients at this column: dTdz(i,j,k-1), dTdz(i,j,k)
    const double mWc1 = hasKp ? (double)a.maskW[c] : 0.0;
    const double mWc0 = hasKm ? (double)a.maskW[c - 1] : 0.0;
    const double dTz_c1 = mWc1 * (Tkp - Tc) * rdz1, dSz_c1 = mWc1 * (Skp - Sc) * rdz1;
    const double dTz_c0 = mWc0 * (Tc - Tkm) * rdz0, dSz_c0 = mWc0 * (Sc - Skm) * rdz0;

    // drdT / drdS at this cell
    double drT_c, drS_c;
    if (Eos<EOS>::kExpensive) {
        drT_c = __ldg(a.drdT + c);
        drS_c = __ldg(a.drdS + c);
    } else {
        eos_drho<EOS>(Sc, Tc, pk, drT_c, drS_c);
        const double m = (double)a.maskT[c];
        drT_c *= m;
        drS_c *= m;
    }

    const TaperParams tp = {a.iso_slopec / a.iso_dslope, 1.0 / a.iso_dslope};
    const double rdzt4 = 1.0 / (4.0 * __ldg(a.g.dzt + k));

    // ---- east face: Ai_ez, K_11 (isoneutral.py:100-132) ---------------------------------------
    double Tke = 0.0, Ske = 0.0, Tkpe = 0.0, Skpe = 0.0;  // (i+1,j,k), (i+1,j,k+1): reused by the top face
    if (inE || inT) {
        Tke = ld(T, ce);
        Ske = ld(S, ce);
        if (hasKp) {
            Tkpe = ld(T, ce + 1);
            Skpe = ld(S, ce + 1);
        }
    }
    const double rdx_c = 1.0 / (__ldg(a.g.dxu + i) * __ldg(a.g.cost + j));
    const double mU_c = (double)a.maskU[c];
    const double dTx_c = mU_c * (Tke - Tc) * rdx_c, dSx_c = mU_c * (Ske - Sc) * rdx_c;
    if (inE) {
        const double Tkme = hasKm ? ld(T, ce - 1) : 0.0, Skme = hasKm ? ld(S, ce - 1) : 0.0;
        const double mWe1 = hasKp ? (double)a.maskW[ce] : 0.0;
        const double mWe0 = hasKm ? (double)a.maskW[ce - 1] : 0.0;
        const double dTz_e1 = mWe1 * (Tkpe - Tke) * rdz1, dSz_e1 = mWe1 * (Skpe - Ske) * rdz1;
        const double dTz_e0 = mWe0 * (Tke - Tkme) * rdz0, dSz_e0 = mWe0 * (Ske - Skme) * rdz0;
        double drT_e, drS_e;
        if (Eos<EOS>::kExpensive) {
            drT_e = __ldg(a.drdT + ce);
            drS_e = __ldg(a.drdS + ce);
        } else {
            eos_drho<EOS>(Ske, Tke, pk, drT_e, drS_e);
            const double m = (double)a.maskT[ce];
            drT_e *= m;
            drS_e *= m;
        }
        double diffloc;
        if (hasKm)
            diffloc = 0.25 * (__ldg(a.K_iso + c) + __ldg(a.K_iso + c - 1) + __ldg(a.K_iso + ce) + __ldg(a.K_iso + ce - 1));
        else
            diffloc = 0.5 * (__ldg(a.K_iso + c) + __ldg(a.K_iso + ce));

        double A[2][2];  // [ip][kr]
        double sumz = 0.0;
#pragma unroll
        for (int kr = 0; kr < 2; ++kr) {
#pragma unroll
            for (int ip = 0; ip < 2; ++ip) {
                const double dT_ = ip ? drT_e : drT_c, dS_ = ip ? drS_e : drS_c;
                const double tz = ip ? (kr ? dTz_e1 : dTz_e0) : (kr ? dTz_c1 : dTz_c0);
                const double sz = ip ? (kr ? dSz_e1 : dSz_e0) : (kr ? dSz_c1 : dSz_c0);
                const double drodxe = dT_ * dTx_c + dS_ * dSx_c;
                const double drodze = dT_ * tz + dS_ * sz;
                const double sxe = -drodxe / (fmin(0.0, drodze) - kEps);
                const double taper = dm_taper(sxe, tp);
                const double dz = kr ? dzw1 : dzw0;
                if (kr == 1 || hasKm) sumz += dz * mU_c * fmax(a.K_iso_steep, diffloc * taper);
                A[ip][kr] = taper * sxe * mU_c;
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
        a.K_11[c] = sumz * rdzt4;
    }

    // ---- north face: Ai_nz, K_22 (isoneutral.py:137-168) --------------------------------------
    double Tkn = 0.0, Skn = 0.0, Tkpn = 0.0, Skpn = 0.0;  // (i,j+1,k), (i,j+1,k+1)
    if (inN || inT) {
        Tkn = ld(T, cn);
        Skn = ld(S, cn);
        if (hasKp) {
            Tkpn = ld(T, cn + 1);
            Skpn = ld(S, cn + 1);
        }
    }
    const double rdy_c = 1.0 / __ldg(a.g.dyu + j);
    const double mV_c = (double)a.maskV[c];
    const double dTy_c = mV_c * (Tkn - Tc) * rdy_c, dSy_c = mV_c * (Skn - Sc) * rdy_c;
    if (inN) {
        const double Tkmn = hasKm ? ld(T, cn - 1) : 0.0, Skmn = hasKm ? ld(S, cn - 1) : 0.0;
        const double mWn1 = hasKp ? (double)a.maskW[cn] : 0.0;
        const double mWn0 = hasKm ? (double)a.maskW[cn - 1] : 0.0;
        const double dTz_n1 = mWn1 * (Tkpn - Tkn) * rdz1, dSz_n1 = mWn1 * (Skpn - Skn) * rdz1;
        const double dTz_n0 = mWn0 * (Tkn - Tkmn) * rdz0, dSz_n0 = mWn0 * (Skn - Skmn) * rdz0;
        double drT_n, drS_n;
        if (Eos<EOS>::kExpensive) {
            drT_n = __ldg(a.drdT + cn);
            drS_n = __ldg(a.drdS + cn);
        } else {
            eos_drho<EOS>(Skn, Tkn, pk, drT_n, drS_n);
            const double m = (double)a.maskT[cn];
            drT_n *= m;
            drS_n *= m;
        }
        double diffloc;
        if (hasKm)
            diffloc = 0.25 * (__ldg(a.K_iso + c) + __ldg(a.K_iso + c - 1) + __ldg(a.K_iso + cn) + __ldg(a.K_iso + cn - 1));
        else
            diffloc = 0.5 * (__ldg(a.K_iso + c) + __ldg(a.K_iso + cn));

        double A[2][2];  // [jp][kr]
        double sumz = 0.0;
#pragma unroll
        for (int kr = 0; kr < 2; ++kr) {
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                const double dT_ = jp ? drT_n : drT_c, dS_ = jp ? drS_n : drS_c;
                const double tz = jp ? (kr ? dTz_n1 : dTz_n0) : (kr ? dTz_c1 : dTz_c0);
                const double sz = jp ? (kr ? dSz_n1 : dSz_n0) : (kr ? dSz_c1 : dSz_c0);
                const double drodyn = dT_ * dTy_c + dS_ * dSy_c;
                const double drodzn = dT_ * tz + dS_ * sz;
                const double syn = -drodyn / (fmin(0.0, drodzn) - kEps);
                const double taper = dm_taper(syn, tp);
                const double dz = kr ? dzw1 : dzw0;
                if (kr == 1 || hasKm) sumz += dz * mV_c * fmax(a.K_iso_steep, diffloc * taper);
                A[jp][kr] = taper * syn * mV_c;
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
        a.K_22[c] = sumz * rdzt4;
    }

    // ---- top face: Ai_bx, Ai_by, K_33 (isoneutral.py:173-225) ---------------------------------
    if (inT) {
        // x and y gradients on the faces around (i,j) at levels k and k+1
        const double rdx_w = 1.0 / (__ldg(a.g.dxu + i - 1) * __ldg(a.g.cost + j));
        const double rdy_s = 1.0 / __ldg(a.g.dyu + j - 1);
        const double Tw = ld(T, cw), Sw = ld(S, cw), Tkpw = ld(T, cw + 1), Skpw = ld(S, cw + 1);
        const double Ts = ld(T, cs), Ss = ld(S, cs), Tkps = ld(T, cs + 1), Skps = ld(S, cs + 1);
        const double mU_cu = (double)a.maskU[c + 1], mU_w = (double)a.maskU[cw], mU_wu = (double)a.maskU[cw + 1];
        const double mV_cu = (double)a.maskV[c + 1], mV_s = (double)a.maskV[cs], mV_su = (double)a.maskV[cs + 1];
        // [ip or jp][kr]
        const double dTx[2][2] = {{mU_w * (Tc - Tw) * rdx_w, mU_wu * (Tkp - Tkpw) * rdx_w},
                                  {dTx_c, mU_cu * (Tkpe - Tkp) * rdx_c}};
        const double dSx[2][2] = {{mU_w * (Sc - Sw) * rdx_w, mU_wu * (Skp - Skpw) * rdx_w},
                                  {dSx_c, mU_cu * (Skpe - Skp) * rdx_c}};
        const double dTy[2][2] = {{mV_s * (Tc - Ts) * rdy_s, mV_su * (Tkp - Tkps) * rdy_s},
                                  {dTy_c, mV_cu * (Tkpn - Tkp) * rdy_c}};
        const double dSy[2][2] = {{mV_s * (Sc - Ss) * rdy_s, mV_su * (Skp - Skps) * rdy_s},
                                  {dSy_c, mV_cu * (Skpn - Skp) * rdy_c}};
        double drT_u, drS_u;  // (i,j,k+1)
        if (Eos<EOS>::kExpensive) {
            drT_u = __ldg(a.drdT + c + 1);
            drS_u = __ldg(a.drdS + c + 1);
        } else {
            eos_drho<EOS>(Skp, Tkp, fabs(__ldg(a.g.zt + k + 1)), drT_u, drS_u);
            const double m = (double)a.maskT[c + 1];
            drT_u *= m;
            drS_u *= m;
        }
        const double Kc = __ldg(a.K_iso + c);
        const double dxu_[2] = {__ldg(a.g.dxu + i - 1), __ldg(a.g.dxu + i)};
        const double facty[2] = {__ldg(a.g.cosu + j - 1) * __ldg(a.g.dyu + j - 1),
                                 __ldg(a.g.cosu + j) * __ldg(a.g.dyu + j)};
        double Ax[2][2], Ay[2][2];
        double sumx = 0.0, sumy = 0.0;
#pragma unroll
        for (int kr = 0; kr < 2; ++kr) {
            const double dT_ = kr ? drT_u : drT_c, dS_ = kr ? drS_u : drS_c;
            const double drodzb = dT_ * dTz_c1 + dS_ * dSz_c1;
            const double rden = 1.0 / (fmin(0.0, drodzb) - kEps);
#pragma unroll
            for (int ip = 0; ip < 2; ++ip) {
                const double drodxb = dT_ * dTx[ip][kr] + dS_ * dSx[ip][kr];
                const double sxb = -drodxb * rden;
                const double taper = dm_taper(sxb, tp);
                sumx += dxu_[ip] * Kc * taper * (sxb * sxb) * mWc1;
                Ax[ip][kr] = taper * sxb * mWc1;
            }
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                const double drodyb = dT_ * dTy[jp][kr] + dS_ * dSy[jp][kr];
                const double syb = -drodyb * rden;
                const double taper = dm_taper(syb, tp);
                sumy += facty[jp] * Kc * taper * (syb * syb) * mWc1;
                Ay[jp][kr] = taper * syb * mWc1;
            }
        }
        store_pair(a.Ai_bx + c * 4, Ax[0][0], Ax[0][1]);
        store_pair(a.Ai_bx + c * 4 + 2, Ax[1][0], Ax[1][1]);
        store_pair(a.Ai_by + c * 4, Ay[0][0], Ay[0][1]);
        store_pair(a.Ai_by + c * 4 + 2, Ay[1][0], Ay[1][1]);
        a.K_33[c] = sumx / (4.0 * __ldg(a.g.dxt + i)) +
                    sumy / (4.0 * __ldg(a.g.dyt + j) * __ldg(a.g.cost + j));
    }
}

void launch_iso_pre(cudaStream_t s, const PreArgs& a) {
    const int N = a.g.N, M = a.g.M, nz = a.g.nz;
    const size_t ncell = (size_t)N * M * nz;
    if (ncell == 0) return;
    if (a.eos == 5) {
        eos5_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, s>>>(ncell, nz, a.temp, a.salt, a.tau, a.maskT, a.g.zt,
                                                                   a.drdT, a.drdS);
        count_launch();
        if (!check_launch("eos5_kernel")) return;
    }
    const int block = 128;
    dim3 grid((M * nz + block - 1) / block, N);
    switch (a.eos) {
    case 1: iso_pre_kernel<1><<<grid, block, 0, s>>>(a); break;
    case 2: iso_pre_kernel<2><<<grid, block, 0, s>>>(a); break;
    case 3: iso_pre_kernel<3><<<grid, block, 0, s>>>(a); break;
    case 4: iso_pre_kernel<4><<<grid, block, 0, s>>>(a); break;
    default: iso_pre_kernel<5><<<grid, block, 0, s>>>(a); break;
    }
    count_launch();
    check_launch("iso_pre_kernel");
}

}  // namespace vb
