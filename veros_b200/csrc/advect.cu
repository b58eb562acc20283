// Tracer advection and the Adams-Bashforth step -- the producer of temp/salt[..., taup1] right before the isoneutral
// path:
//   advect_tracer            veros/core/thermodynamics.py:10-40
//   adv_flux_2nd / superbee  veros/core/advection.py:8-115 (flux limiter of :8-20, pad_z_edges in the vertical)
//   advect_temperature / advect_salinity   thermodynamics.py:43-62   (d{temp,salt}[..., tau] = advect_tracer(...))
//   Adams-Bashforth step     thermodynamics.py:223-245
// One thread per cell computes the six face fluxes around it (each face is evaluated by the two cells that share it:
// 2x the arithmetic of the reference's three flux arrays, none of their 48 B/cell of HBM traffic), the tendency and
// the new time level of both tracers in one pass.  Every operation is an explicitly rounded intrinsic in the
// reference's order: bit-identical to the NumPy backend.
#include "common.cuh"
#include "strict.cuh"

namespace vb {
namespace {

using strict::add;
using strict::mul;
using strict::sub;

struct AdvArgs {
    int N, M, nz;
    double* tr[2];    // temp, salt (N,M,nz,3): level tau read, taup1 written
    double* dtr[2];   // dtemp, dsalt (N,M,nz,3): level tau written, taum1 read
    const int32_t *tau, *taup1, *taum1;
    const double *u, *v, *w;  // (N,M,nz,3)
    const uint8_t *maskT, *maskU, *maskV, *maskW;
    const double *dxt, *dyt, *dzt, *cost, *cosu;
    double dt, c1, c2;  // dt_tracer, 1.5 + AB_eps, 0.5 + AB_eps
    int superbee, with_ab;
};

__device__ __forceinline__ double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// advection.py:39-49 for one face
__device__ __forceinline__ double superbee_face(double vm1, double v0, double v1, double v2, double mm1, double m0, double m1,
                                                double vel, double velc, double dt, double dx) {
    const double eps = 1e-20;
    const double rjp = mul(sub(v2, v1), m1), rj = mul(sub(v1, v0), m0), rjm = mul(sub(v0, vm1), mm1);
    const double num = vel > 0.0 ? rjm : rjp;
    const double den = fabs(rj) < eps ? eps : rj;
    const double c = __ddiv_rn(num, den);
    const double a1 = clipd(mul(2.0, c), 0.0, 1.0), a2 = clipd(c, 0.0, 2.0);
    const double cr = a1 > a2 ? a1 : a2;
    const double uCFL = fabs(__ddiv_rn(mul(velc, dt), dx));
    return sub(mul(mul(velc, add(v1, v0)), 0.5), mul(mul(mul(fabs(velc), add(sub(1.0, cr), mul(uCFL, cr))), rj), 0.5));
}

template <bool SUPERBEE>
__global__ void __launch_bounds__(256)
advect_kernel(const AdvArgs a) {
    const int N = a.N, M = a.M, nz = a.nz;
    const size_t n3 = (size_t)N * M * nz;
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n3) return;
    const size_t plane = (size_t)M * nz;
    const int i = (int)(c / plane);
    const int j = (int)((c - (size_t)i * plane) / nz);
    const int k = (int)(c % nz);
    const int tau = *a.tau, taup1 = *a.taup1, taum1 = *a.taum1;
    const bool interior = i >= 2 && i < N - 2 && j >= 2 && j < M - 2;
    const double mT = (double)a.maskT[c];

    // velocities and masks of the faces around the cell (shared by both tracers)
    double ue = 0, uw = 0, vn = 0, vs_ = 0, wt = 0, wb = 0;
    double mU[4] = {0, 0, 0, 0}, mV[4] = {0, 0, 0, 0}, mW[4] = {0, 0, 0, 0};  // masks at offsets -2 .. +1 of the cell
    double cdx = 1.0, cdy = 1.0, cdyn = 1.0, cdys = 1.0, cun = 0.0, cus = 0.0;
    const int km1 = max(k - 1, 0), km2 = max(k - 2, 0), kp1 = min(k + 1, nz - 1), kp2 = min(k + 2, nz - 1);
    if (interior) {
        ue = a.u[c * 3 + tau];
        uw = a.u[(c - plane) * 3 + tau];
        vn = a.v[c * 3 + tau];
        vs_ = a.v[(c - nz) * 3 + tau];
        wt = k < nz - 1 ? a.w[c * 3 + tau] : 0.0;
        wb = k > 0 ? a.w[(c - 1) * 3 + tau] : 0.0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            mU[o] = (double)a.maskU[c + (ptrdiff_t)(o - 2) * (ptrdiff_t)plane];
            mV[o] = (double)a.maskV[c + (ptrdiff_t)(o - 2) * nz];
        }
        const size_t col = c - k;
        mW[0] = (double)a.maskW[col + km2];
        mW[1] = (double)a.maskW[col + km1];
        mW[2] = (double)a.maskW[c];
        mW[3] = (double)a.maskW[col + kp1];
        cdx = mul(a.cost[j], a.dxt[i]);            // cost[j] * dxt[i]: superbee dx of the east face AND the divisor
        cdy = mul(a.cost[j], a.dyt[j]);
        cdyn = cdy;                                // (cost * dyt)[j]
        cdys = mul(a.cost[j - 1], a.dyt[j - 1]);
        cun = a.cosu[j];
        cus = a.cosu[j - 1];
    }
    const double cdxw = interior ? mul(a.cost[j], a.dxt[i - 1]) : 1.0;

#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const double* __restrict__ tr = a.tr[t] + tau;
        const double T0 = tr[c * 3];
        double d = 0.0;
        if (interior) {
            auto X = [&](int o) { return tr[(c + (ptrdiff_t)o * (ptrdiff_t)plane) * 3]; };
            auto Y = [&](int o) { return tr[(c + (ptrdiff_t)o * nz) * 3]; };
            const size_t col = c - k;
            const double Tw2 = X(-2), Tw = X(-1), Te = X(1), Te2 = X(2);
            const double Ts2 = Y(-2), Ts = Y(-1), Tn = Y(1), Tn2 = Y(2);
            const double Tb2 = tr[(col + km2) * 3], Tb = tr[(col + km1) * 3], Tt = tr[(col + kp1) * 3], Tt2 = tr[(col + kp2) * 3];
            double fe, fw, fn, fs, ft = 0.0, fb = 0.0;
            if (SUPERBEE) {
                fe = superbee_face(Tw, T0, Te, Te2, mU[1], mU[2], mU[3], ue, ue, a.dt, cdx);
                fw = superbee_face(Tw2, Tw, T0, Te, mU[0], mU[1], mU[2], uw, uw, a.dt, cdxw);
                fn = superbee_face(Ts, T0, Tn, Tn2, mV[1], mV[2], mV[3], vn, mul(vn, cun), a.dt, cdyn);
                fs = superbee_face(Ts2, Ts, T0, Tn, mV[0], mV[1], mV[2], vs_, mul(vs_, cus), a.dt, cdys);
                // vertical: pad_z_edges clamps the outermost neighbours (advection.py:36-37)
                if (k < nz - 1) ft = superbee_face(Tb, T0, Tt, Tt2, mW[1], mW[2], mW[3], wt, wt, a.dt, a.dzt[k]);
                if (k > 0) fb = superbee_face(Tb2, Tb, T0, Tt, mW[0], mW[1], mW[2], wb, wb, a.dt, a.dzt[k - 1]);
            } else {
                fe = mul(mul(mul(0.5, add(T0, Te)), ue), mU[2]);
                fw = mul(mul(mul(0.5, add(Tw, T0)), uw), mU[1]);
                fn = mul(mul(mul(mul(cun, 0.5), add(T0, Tn)), vn), mV[2]);
                fs = mul(mul(mul(mul(cus, 0.5), add(Ts, T0)), vs_), mV[1]);
                if (k < nz - 1) ft = mul(mul(mul(0.5, add(T0, Tt)), wt), mW[2]);
                if (k > 0) fb = mul(mul(mul(0.5, add(Tb, T0)), wb), mW[1]);
            }
            // thermodynamics.py:24-39
            d = mul(mT, sub(__ddiv_rn(-sub(fe, fw), cdx), __ddiv_rn(sub(fn, fs), cdy)));
            const double mt = mT != 0.0 ? -1.0 : 0.0;  // -1 * maskT, an integer product in the reference: never -0
            if (k == 0)
                d = add(d, __ddiv_rn(mul(mt, ft), a.dzt[0]));
            else
                d = add(d, __ddiv_rn(mul(mt, sub(ft, fb)), a.dzt[k]));
        }
        a.dtr[t][c * 3 + tau] = d;
        if (a.with_ab) {  // thermodynamics.py:226-243, every cell of the array
            const double dm1 = a.dtr[t][c * 3 + taum1];
            a.tr[t][c * 3 + taup1] = add(T0, mul(mul(a.dt, sub(mul(a.c1, d), mul(a.c2, dm1))), mT));
        }
    }
}

}  // namespace

void launch_advect_tempsalt(cudaStream_t s, const VerosB200AdvectDescriptor* d, void** B) {
    AdvArgs a;
    a.N = d->nx_tot;
    a.M = d->ny_tot;
    a.nz = d->nz;
    a.tr[0] = (double*)B[19];
    a.tr[1] = (double*)B[20];
    a.dtr[0] = (double*)B[21];
    a.dtr[1] = (double*)B[22];
    a.tau = (const int32_t*)B[4];
    a.taup1 = (const int32_t*)B[5];
    a.taum1 = (const int32_t*)B[6];
    a.u = (const double*)B[7];
    a.v = (const double*)B[8];
    a.w = (const double*)B[9];
    a.maskT = (const uint8_t*)B[10];
    a.maskU = (const uint8_t*)B[11];
    a.maskV = (const uint8_t*)B[12];
    a.maskW = (const uint8_t*)B[13];
    a.dxt = (const double*)B[14];
    a.dyt = (const double*)B[15];
    a.dzt = (const double*)B[16];
    a.cost = (const double*)B[17];
    a.cosu = (const double*)B[18];
    a.dt = d->dt_tracer;
    a.c1 = 1.5 + d->AB_eps;
    a.c2 = 0.5 + d->AB_eps;
    a.superbee = (d->flags & VEROS_B200_ADVECT_SUPERBEE) != 0;
    a.with_ab = (d->flags & VEROS_B200_ADVECT_NO_AB) == 0;
    const size_t n3 = (size_t)a.N * a.M * a.nz;
    if (n3 == 0) return;
    const unsigned grid = (unsigned)((n3 + 255) / 256);
    if (a.superbee)
        advect_kernel<true><<<grid, 256, 0, s>>>(a);
    else
        advect_kernel<false><<<grid, 256, 0, s>>>(a);
    count_launch();
    check_launch("advect_kernel");
}

}  // namespace vb
