/*
 * oracle/iso_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement (plain C, IEEE double, no FMA contraction; optional OpenMP over x-planes,
 * which does not change any result bit) of the
 * reference's isoneutral-mixing hot path.  It exists only to CHECK the CUDA kernels:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (veros_b200/) never imports, links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * golden vectors produced by the reference's own NumPy implementation (imported from
 * /root/reference by tests/golden/make_golden.py, fixtures committed under tests/golden/):
 *   - isoneutral_diffusion (T, S, skew), solve_tridiagonal: bit-for-bit (operation order of the
 *     reference's NumPy expressions and of LAPACK dgtsv's no-pivot elimination is mirrored);
 *   - isoneutral_diffusion_pre: <= 2e-15 normalised (libm tanh vs NumPy's SIMD tanh).
 *
 * Reference files followed (paths relative to the reference checkout):
 *   veros/core/isoneutral/isoneutral.py:10-229   dm_taper, isoneutral_diffusion_pre
 *   veros/core/isoneutral/diffusion.py:9-283     fluxes, explicit, implicit, tracer update, energy
 *   veros/core/utilities.py:24-59                pad_z_edges, create_water_masks, solve_implicit
 *   veros/core/operators.py:60-77                solve_tridiagonal_numpy (LAPACK dgtsv)
 *   veros/core/operators.py:133-156, special/tdma_cython_.pyx:8-25   Thomas cp/dp recurrence
 *   veros/core/diffusion.py:9-62                 compute_dissipation, dissipation_on_wgrid
 *   veros/core/thermodynamics.py:248-300         vertmix_tempsalt (golden vectors
 *                                                tests/golden/vmix_*.npz, bit-for-bit)
 *   veros/core/utilities.py:8-20                 enforce_boundaries
 *   veros/core/density/{linear_eq,nonlinear_eq1,nonlinear_eq2,nonlinear_eq3,gsw}.py, get_rho.py:93-131
 *
 * Layout: C order, z fastest: f[i][j][k] -> (i*M + j)*nz + k ; tracers (N,M,nz,3) time level last;
 * Ai_* (N,M,nz,2,2) -> ((cell*2)+ip)*2+kr.  Index 0 in z is the bottom.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t N, M, nz;               /* sizes including the 2+2 ghost cells in x and y */
    int32_t eos_type;               /* settings.eq_of_state_type 1..5 */
    int32_t enable_conserve_energy; /* settings.enable_conserve_energy */
    int32_t tau, taup1;             /* time-level indices into the (..,3) tracer arrays */
    int32_t pad_;
    double K_iso_steep, iso_slopec, iso_dslope, dt_tracer, grav, rho_0;
} oracle_params;

#define IDX(i, j, k) ((((size_t)(i)) * M + (j)) * nz + (k))
#define TIDX(i, j, k, t) (IDX(i, j, k) * 3 + (t))
#define AIDX(i, j, k, ip, kr) ((IDX(i, j, k) * 2 + (ip)) * 2 + (kr))

/* ------------------------------------------------------------------------------------------
 * Equation of state derivatives (veros/core/density/ *.py files), evaluated at p = |zt[k]|
 * ---------------------------------------------------------------------------------------- */
static const double v01 = 9.998420897506056e2, v02 = 2.839940833161907e0, v03 = -3.147759265588511e-2,
                    v04 = 1.181805545074306e-3, v05 = -6.698001071123802e0, v06 = -2.986498947203215e-2,
                    v07 = 2.327859407479162e-4, v08 = -3.988822378968490e-2, v09 = 5.095422573880500e-4,
                    v10 = -1.426984671633621e-5, v11 = 1.645039373682922e-7, v12 = -2.233269627352527e-2,
                    v13 = -3.436090079851880e-4, v14 = 3.726050720345733e-6, v15 = -1.806789763745328e-4,
                    v16 = 6.876837219536232e-7, v17 = -3.087032500374211e-7, v18 = -1.988366587925593e-8,
                    v19 = -1.061519070296458e-11, v20 = 1.550932729220080e-10, v21 = 1.0e0,
                    v22 = 2.775927747785646e-3, v23 = -2.349607444135925e-5, v24 = 1.119513357486743e-6,
                    v25 = 6.743689325042773e-10, v26 = -7.521448093615448e-3, v27 = -2.764306979894411e-5,
                    v28 = 1.262937315098546e-7, v29 = 9.527875081696435e-10, v30 = -1.811147201949891e-11,
                    v31 = -3.303308871386421e-5, v32 = 3.801564588876298e-7, v33 = -7.672876869259043e-9,
                    v34 = -4.634182341116144e-11, v35 = 2.681097235569143e-12, v36 = 5.419326551148740e-6,
                    v37 = -2.742185394906099e-5, v38 = -3.212746477974189e-7, v39 = 3.191413910561627e-9,
                    v40 = -1.931012931541776e-12, v41 = -1.105097577149576e-7, v42 = 6.211426728363857e-10,
                    v43 = -1.119011592875110e-10, v44 = -1.941660213148725e-11, v45 = -1.864826425365600e-14,
                    v46 = 1.119522344879478e-14, v47 = -1.200507748551599e-15, v48 = 6.057902487546866e-17;

/* gsw.py:14-63 -- the two 48-term polynomials shared by every gsw derivative */
static void gsw_den_num(double sa, double ct, double p, double sqrtsa, double *den, double *num) {
    *den = v01 + ct * (v02 + ct * (v03 + v04 * ct)) +
           sa * (v05 + ct * (v06 + v07 * ct) + sqrtsa * (v08 + ct * (v09 + ct * (v10 + v11 * ct)))) +
           p * (v12 + ct * (v13 + v14 * ct) + sa * (v15 + v16 * ct) +
                p * (v17 + ct * (v18 + v19 * ct) + v20 * sa));
    *num = v21 + ct * (v22 + ct * (v23 + ct * (v24 + v25 * ct))) +
           sa * (v26 + ct * (v27 + ct * (v28 + ct * (v29 + v30 * ct))) + v36 * sa +
                 sqrtsa * (v31 + ct * (v32 + ct * (v33 + ct * (v34 + v35 * ct))))) +
           p * (v37 + ct * (v38 + ct * (v39 + v40 * ct)) + sa * (v41 + v42 * ct) +
                p * (v43 + ct * (v44 + v45 * ct + v46 * sa) + p * (v47 + v48 * ct)));
}

/* gsw.py:105-192 */
static double gsw_drhodT(double sa, double ct, double p) {
    const double a01 = 2.839940833161907e0, a02 = -6.295518531177023e-2, a03 = 3.545416635222918e-3,
                 a04 = -2.986498947203215e-2, a05 = 4.655718814958324e-4, a06 = 5.095422573880500e-4,
                 a07 = -2.853969343267241e-5, a08 = 4.935118121048767e-7, a09 = -3.436090079851880e-4,
                 a10 = 7.452101440691467e-6, a11 = 6.876837219536232e-7, a12 = -1.988366587925593e-8,
                 a13 = -2.123038140592916e-11, a14 = 2.775927747785646e-3, a15 = -4.699214888271850e-5,
                 a16 = 3.358540072460230e-6, a17 = 2.697475730017109e-9, a18 = -2.764306979894411e-5,
                 a19 = 2.525874630197091e-7, a20 = 2.858362524508931e-9, a21 = -7.244588807799565e-11,
                 a22 = 3.801564588876298e-7, a23 = -1.534575373851809e-8, a24 = -1.390254702334843e-10,
                 a25 = 1.072438894227657e-11, a26 = -3.212746477974189e-7, a27 = 6.382827821123254e-9,
                 a28 = -5.793038794625329e-12, a29 = 6.211426728363857e-10, a30 = -1.941660213148725e-11,
                 a31 = -3.729652850731201e-14, a32 = 1.119522344879478e-14, a33 = 6.057902487546866e-17;
    double sqrtsa = sqrt(sa), den, num;
    gsw_den_num(sa, ct, p, sqrtsa, &den, &num);
    double dden = a01 + ct * (a02 + a03 * ct) + sa * (a04 + a05 * ct + sqrtsa * (a06 + ct * (a07 + a08 * ct))) +
                  p * (a09 + a10 * ct + a11 * sa + p * (a12 + a13 * ct));
    double dnum = a14 + ct * (a15 + ct * (a16 + a17 * ct)) +
                  sa * (a18 + ct * (a19 + ct * (a20 + a21 * ct)) +
                        sqrtsa * (a22 + ct * (a23 + ct * (a24 + a25 * ct)))) +
                  p * (a26 + ct * (a27 + a28 * ct) + a29 * sa + p * (a30 + a31 * ct + a32 * sa + a33 * p));
    double rec_num = 1.0 / num;
    double rho = rec_num * den;
    return (dden - dnum * rho) * rec_num;
}

/* gsw.py:195-274 */
static double gsw_drhodS(double sa, double ct, double p) {
    const double b01 = -6.698001071123802e0, b02 = -2.986498947203215e-2, b03 = 2.327859407479162e-4,
                 b04 = -5.983233568452735e-2, b05 = 7.643133860820750e-4, b06 = -2.140477007450431e-5,
                 b07 = 2.467559060524383e-7, b08 = -1.806789763745328e-4, b09 = 6.876837219536232e-7,
                 b10 = 1.550932729220080e-10, b11 = -7.521448093615448e-3, b12 = -2.764306979894411e-5,
                 b13 = 1.262937315098546e-7, b14 = 9.527875081696435e-10, b15 = -1.811147201949891e-11,
                 b16 = -4.954963307079632e-5, b17 = 5.702346883314446e-7, b18 = -1.150931530388857e-8,
                 b19 = -6.951273511674217e-11, b20 = 4.021645853353715e-12, b21 = 1.083865310229748e-5,
                 b22 = -1.105097577149576e-7, b23 = 6.211426728363857e-10, b24 = 1.119522344879478e-14;
    double sqrtsa = sqrt(sa), den, num;
    gsw_den_num(sa, ct, p, sqrtsa, &den, &num);
    double dden = b01 + ct * (b02 + b03 * ct) + sqrtsa * (b04 + ct * (b05 + ct * (b06 + b07 * ct))) +
                  p * (b08 + b09 * ct + b10 * p);
    double dnum = b11 + ct * (b12 + ct * (b13 + ct * (b14 + b15 * ct))) +
                  sqrtsa * (b16 + ct * (b17 + ct * (b18 + ct * (b19 + b20 * ct)))) + b21 * sa +
                  p * (b22 + ct * (b23 + b24 * p));
    double rec_num = 1.0 / num;
    double rho = rec_num * den;
    return (dden - dnum * rho) * rec_num;
}

/* get_rho.py:93-131 dispatch; bodies linear_eq.py:37-43, nonlinear_eq1.py:40-48,
 * nonlinear_eq2.py:58-66, nonlinear_eq3.py:34-41 */
static double eos_drhodT(int type, double sa, double ct, double p) {
    const double rho0 = 1024.0, theta0 = 283.0 - 273.15, betaT = 1.67e-4, grav = 9.81, z0 = 0.0;
    switch (type) {
    case 1: return -betaT * rho0;
    case 2:
    case 4: {
        const double betaTs = 1e-5 / 2.0;
        double thetas = ct - theta0;
        return -(betaT + 2 * betaTs * thetas) * rho0;
    }
    case 3: {
        const double betaTs = 1e-5, gammas = 1.1e-8;
        double zz = -p - z0;
        double thetas = ct - theta0;
        return -(betaTs * thetas + betaT * (1 - gammas * grav * zz * rho0)) * rho0;
    }
    default: return gsw_drhodT(sa, ct, p);
    }
}

static double eos_drhodS(int type, double sa, double ct, double p) {
    const double rho0 = 1024.0;
    switch (type) {
    case 1:
    case 2:
    case 3: return 0.78e-3 * rho0;
    case 4: return 0 * rho0;
    default: return gsw_drhodS(sa, ct, p);
    }
}

/* isoneutral.py:10-15 */
static double dm_taper(double sx, double slopec, double dslope) {
    return 0.5 * (1.0 + tanh((-fabs(sx) + slopec) / dslope));
}

/* ------------------------------------------------------------------------------------------
 * isoneutral_diffusion_pre   (isoneutral.py:18-229)
 * Ai_* / K_* are in/out: only the reference's write regions are touched.
 * ---------------------------------------------------------------------------------------- */
void oracle_iso_pre(const oracle_params *P, const double *temp, const double *salt, const double *K_iso,
                    const uint8_t *maskT, const uint8_t *maskU, const uint8_t *maskV, const uint8_t *maskW,
                    const double *dxt, const double *dxu, const double *dyt, const double *dyu,
                    const double *cost, const double *cosu, const double *dzt, const double *dzw,
                    const double *zt, double *Ai_ez, double *Ai_nz, double *Ai_bx, double *Ai_by,
                    double *K_11, double *K_22, double *K_33) {
    const int N = P->N, M = P->M, nz = P->nz, tau = P->tau;
    const double epsln = 1e-20;
    const size_t n3 = (size_t)N * M * nz;
    double *drdT = calloc(n3, 8), *drdS = calloc(n3, 8);
    double *dTdx = calloc(n3, 8), *dSdx = calloc(n3, 8), *dTdy = calloc(n3, 8), *dSdy = calloc(n3, 8);
    double *dTdz = calloc(n3, 8), *dSdz = calloc(n3, 8);
    (void)dxt;

#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++)
        for (int j = 0; j < M; j++)
            for (int k = 0; k < nz; k++) {
                double T = temp[TIDX(i, j, k, tau)], S = salt[TIDX(i, j, k, tau)];
                double p = fabs(zt[k]);
                double mT = (double)maskT[IDX(i, j, k)];
                drdT[IDX(i, j, k)] = mT * eos_drhodT(P->eos_type, S, T, p); /* :40 */
                drdS[IDX(i, j, k)] = mT * eos_drhodS(P->eos_type, S, T, p); /* :41 */
                if (k < nz - 1) {                                           /* :46-59 */
                    double mW = (double)maskW[IDX(i, j, k)];
                    dTdz[IDX(i, j, k)] = mW * (temp[TIDX(i, j, k + 1, tau)] - T) / dzw[k];
                    dSdz[IDX(i, j, k)] = mW * (salt[TIDX(i, j, k + 1, tau)] - S) / dzw[k];
                }
                if (i < N - 1) { /* :64-77 */
                    double mU = (double)maskU[IDX(i, j, k)];
                    dTdx[IDX(i, j, k)] = mU * (temp[TIDX(i + 1, j, k, tau)] - T) / (dxu[i] * cost[j]);
                    dSdx[IDX(i, j, k)] = mU * (salt[TIDX(i + 1, j, k, tau)] - S) / (dxu[i] * cost[j]);
                }
                if (j < M - 1) { /* :82-95 */
                    double mV = (double)maskV[IDX(i, j, k)];
                    dTdy[IDX(i, j, k)] = mV * (temp[TIDX(i, j + 1, k, tau)] - T) / dyu[j];
                    dSdy[IDX(i, j, k)] = mV * (salt[TIDX(i, j + 1, k, tau)] - S) / dyu[j];
                }
            }

    /* east face: Ai_ez, K_11  (:100-132) */
#pragma omp parallel for schedule(static)
    for (int i = 1; i < N - 2; i++)
        for (int j = 2; j < M - 2; j++)
            for (int k = 0; k < nz; k++) {
                double diffloc;
                if (k >= 1)
                    diffloc = 0.25 * (K_iso[IDX(i, j, k)] + K_iso[IDX(i, j, k - 1)] + K_iso[IDX(i + 1, j, k)] +
                                      K_iso[IDX(i + 1, j, k - 1)]);
                else
                    diffloc = 0.5 * (K_iso[IDX(i, j, 0)] + K_iso[IDX(i + 1, j, 0)]);
                double mU = (double)maskU[IDX(i, j, k)];
                double sumz = 0.0;
                for (int kr = 0; kr < 2; kr++) {
                    if (kr == 0 && k == 0) continue; /* ki = 1 for kr = 0 */
                    for (int ip = 0; ip < 2; ip++) {
                        size_t c = IDX(i + ip, j, k), cz = IDX(i + ip, j, k + kr - 1);
                        double drodxe = drdT[c] * dTdx[IDX(i, j, k)] + drdS[c] * dSdx[IDX(i, j, k)];
                        double drodze = drdT[c] * dTdz[cz] + drdS[c] * dSdz[cz];
                        double sxe = -drodxe / (fmin(0.0, drodze) - epsln);
                        double taper = dm_taper(sxe, P->iso_slopec, P->iso_dslope);
                        sumz += dzw[k + kr - 1] * mU * fmax(P->K_iso_steep, diffloc * taper);
                        Ai_ez[AIDX(i, j, k, ip, kr)] = taper * sxe * mU;
                    }
                }
                K_11[IDX(i, j, k)] = sumz / (4.0 * dzt[k]);
            }

    /* north face: Ai_nz, K_22  (:137-168) */
#pragma omp parallel for schedule(static)
    for (int i = 2; i < N - 2; i++)
        for (int j = 1; j < M - 2; j++)
            for (int k = 0; k < nz; k++) {
                double diffloc;
                if (k >= 1)
                    diffloc = 0.25 * (K_iso[IDX(i, j, k)] + K_iso[IDX(i, j, k - 1)] + K_iso[IDX(i, j + 1, k)] +
                                      K_iso[IDX(i, j + 1, k - 1)]);
                else
                    diffloc = 0.5 * (K_iso[IDX(i, j, 0)] + K_iso[IDX(i, j + 1, 0)]);
                double mV = (double)maskV[IDX(i, j, k)];
                double sumz = 0.0;
                for (int kr = 0; kr < 2; kr++) {
                    if (kr == 0 && k == 0) continue;
                    for (int jp = 0; jp < 2; jp++) {
                        size_t c = IDX(i, j + jp, k), cz = IDX(i, j + jp, k + kr - 1);
                        double drodyn = drdT[c] * dTdy[IDX(i, j, k)] + drdS[c] * dSdy[IDX(i, j, k)];
                        double drodzn = drdT[c] * dTdz[cz] + drdS[c] * dSdz[cz];
                        double syn = -drodyn / (fmin(0.0, drodzn) - epsln);
                        double taper = dm_taper(syn, P->iso_slopec, P->iso_dslope);
                        sumz += dzw[k + kr - 1] * mV * fmax(P->K_iso_steep, diffloc * taper);
                        Ai_nz[AIDX(i, j, k, jp, kr)] = taper * syn * mV;
                    }
                }
                K_22[IDX(i, j, k)] = sumz / (4.0 * dzt[k]);
            }

    /* top face: Ai_bx, Ai_by, K_33  (:173-225) */
#pragma omp parallel for schedule(static)
    for (int i = 2; i < N - 2; i++)
        for (int j = 2; j < M - 2; j++)
            for (int k = 0; k < nz - 1; k++) {
                double mW = (double)maskW[IDX(i, j, k)];
                double Kc = K_iso[IDX(i, j, k)];
                double sumx = 0.0, sumy = 0.0;
                for (int kr = 0; kr < 2; kr++) {
                    size_t c = IDX(i, j, k + kr);
                    double drodzb = drdT[c] * dTdz[IDX(i, j, k)] + drdS[c] * dSdz[IDX(i, j, k)];
                    for (int ip = 0; ip < 2; ip++) {
                        size_t cx = IDX(i - 1 + ip, j, k + kr);
                        double drodxb = drdT[c] * dTdx[cx] + drdS[c] * dSdx[cx];
                        double sxb = -drodxb / (fmin(0.0, drodzb) - epsln);
                        double taper = dm_taper(sxb, P->iso_slopec, P->iso_dslope);
                        sumx = sumx + dxu[i - 1 + ip] * Kc * taper * (sxb * sxb) * mW;
                        Ai_bx[AIDX(i, j, k, ip, kr)] = taper * sxb * mW;
                    }
                    for (int jp = 0; jp < 2; jp++) {
                        double facty = cosu[j - 1 + jp] * dyu[j - 1 + jp];
                        size_t cy = IDX(i, j - 1 + jp, k + kr);
                        double drodyb = drdT[c] * dTdy[cy] + drdS[c] * dSdy[cy];
                        double syb = -drodyb / (fmin(0.0, drodzb) - epsln);
                        double taper = dm_taper(syb, P->iso_slopec, P->iso_dslope);
                        sumy = sumy + facty * Kc * taper * (syb * syb) * mW;
                        Ai_by[AIDX(i, j, k, jp, kr)] = taper * syb * mW;
                    }
                }
                K_33[IDX(i, j, k)] = sumx / (4 * dxt[i]) + sumy / (4 * dyt[j] * cost[j]);
            }
    for (int i = 0; i < N; i++) /* :225  K_33[..., -1] = 0 over the whole array */
        for (int j = 0; j < M; j++) K_33[IDX(i, j, nz - 1)] = 0.0;

    free(drdT); free(drdS); free(dTdx); free(dSdx); free(dTdy); free(dSdy); free(dTdz); free(dSdz);
}

/* ------------------------------------------------------------------------------------------
 * Column solve.  mode 0: LAPACK dgtsv operation order, partial pivoting included (what the NumPy
 * backend executes, operators.py:60-77; SciPy's bundled LAPACK, reference pin scipy==1.18.0);
 * mode 1: Thomas cp/dp recurrence with reciprocal (tdma_cython_.pyx:8-25, cuda_tdma_kernels.cu:43-66).
 * Mask semantics of tdma_.py:63-66 / operators.py:133-136: rows outside water are identity rows
 * with zero rhs, the edge row has a = 0.  Result is 0 outside water.
 * ---------------------------------------------------------------------------------------- */
static void solve_column(int nz, const double *a, const double *b, const double *c, const double *d,
                         const uint8_t *water, const uint8_t *edge, double *x, double *w1, double *w2, int mode) {
    for (int k = 0; k < nz; k++) { /* masked copies: w1 = diagonal, w2 = rhs */
        w1[k] = water[k] ? b[k] : 1.0;
        w2[k] = water[k] ? d[k] : 0.0;
    }
#define A_(k) ((water[k] && !edge[k]) ? a[k] : 0.0)
#define C_(k) ((water[k] && (k) < nz - 1) ? c[k] : 0.0)
    if (mode == 0) {
        /* LAPACK dgtsv (NRHS = 1) INCLUDING its partial pivoting, applied to this column alone.
         * In the reference all water cells form one long system (operators.py:75) whose inter-column
         * couplings are zero (a[edge] = 0, c[..., -1] = 0), so no interchange ever crosses a column
         * boundary and the per-column elimination below performs the identical operation sequence.
         * dl[k] couples row k+1 to row k (= a[k+1]); after the sweep dl[k] holds the second
         * super-diagonal created by an interchange (0 otherwise). */
        double *dl = malloc(8 * (size_t)(nz + 2)), *du = malloc(8 * (size_t)(nz + 2));
        for (int k = 0; k < nz; k++) {
            dl[k] = (k + 1 < nz) ? A_(k + 1) : 0.0;
            du[k] = C_(k);
        }
        du[nz] = 0.0;
        for (int k = 0; k < nz - 1; k++) {
            if (fabs(w1[k]) >= fabs(dl[k])) {
                double fact = dl[k] / w1[k];
                w1[k + 1] = w1[k + 1] - fact * du[k];
                w2[k + 1] = w2[k + 1] - fact * w2[k];
                dl[k] = 0.0;
            } else {
                double fact = w1[k] / dl[k];
                w1[k] = dl[k];
                double temp = w1[k + 1];
                w1[k + 1] = du[k] - fact * temp;
                dl[k] = du[k + 1];
                du[k + 1] = -fact * dl[k];
                du[k] = temp;
                temp = w2[k];
                w2[k] = w2[k + 1];
                w2[k + 1] = temp - fact * w2[k + 1];
            }
        }
        x[nz - 1] = w2[nz - 1] / w1[nz - 1];
        if (nz > 1) x[nz - 2] = (w2[nz - 2] - du[nz - 2] * x[nz - 1]) / w1[nz - 2];
        for (int k = nz - 3; k >= 0; k--) x[k] = (w2[k] - du[k] * x[k + 1] - dl[k] * x[k + 2]) / w1[k];
        free(dl); free(du);
    } else {
        double *cp = w1, *dp = w2;
        double b0 = w1[0];
        cp[0] = C_(0) / b0;
        dp[0] = w2[0] / b0;
        for (int k = 1; k < nz; k++) {
            double ak = A_(k);
            double denom = 1.0 / (w1[k] - ak * cp[k - 1]);
            cp[k] = C_(k) * denom;
            dp[k] = (w2[k] - ak * dp[k - 1]) * denom;
        }
        x[nz - 1] = dp[nz - 1];
        for (int k = nz - 2; k >= 0; k--) x[k] = dp[k] - cp[k] * x[k + 1];
    }
#undef A_
#undef C_
    for (int k = 0; k < nz; k++)
        if (!water[k]) x[k] = 0.0;
}

/* utilities.solve_implicit (utilities.py:51-59) + solve_tridiagonal.  b_edge / d_edge may be NULL. */
void oracle_solve_implicit(int64_t ncol, int32_t nz, const double *a, const double *b, const double *c,
                           const double *d, const uint8_t *water, const uint8_t *edge, const double *b_edge,
                           const double *d_edge, double *out, int32_t mode) {
#pragma omp parallel
    {
    double *bb = malloc(8 * (size_t)nz), *dd = malloc(8 * (size_t)nz);
    double *w1 = malloc(8 * (size_t)nz), *w2 = malloc(8 * (size_t)nz);
#pragma omp for schedule(static)
    for (int64_t col = 0; col < ncol; col++) {
        size_t o = (size_t)col * nz;
        for (int k = 0; k < nz; k++) {
            bb[k] = (b_edge && edge[o + k]) ? b_edge[o + k] : b[o + k];
            dd[k] = (d_edge && edge[o + k]) ? d_edge[o + k] : d[o + k];
        }
        solve_column(nz, a + o, bb, c + o, dd, water + o, edge + o, out + o, w1, w2, mode);
    }
    free(bb); free(dd); free(w1); free(w2);
    }
}

/* number of OpenMP threads the loops above will use (1 when built without -fopenmp) */
#ifdef _OPENMP
#include <omp.h>
int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }
#else
int oracle_num_threads(void) { return 1; }
void oracle_set_num_threads(int n) { (void)n; }
#endif

/* ------------------------------------------------------------------------------------------
 * isoneutral_diffusion / isoneutral_skew_diffusion for one tracer
 * (diffusion.py:9-283; routine wrappers :286-307).
 *   iso != 0: K1 = K2 = K_iso, implicit K_33 part, P_diss (= P_diss_iso) gets both terms
 *   iso == 0: K1 = -K_gm, K2 = K_gm (K_iso := K_gm array passed in), no implicit part,
 *             P_diss (= P_diss_skew) gets the explicit vertical term only
 * tr (N,M,nz,3) and dtracer (N,M,nz), P_diss (N,M,nz) are updated in place.
 * flux_east/north/top (N,M,nz) are outputs (zero outside their write regions).
 * int_drhodX (N,M,nz,3) and P_diss may be NULL when enable_conserve_energy == 0.
 * ---------------------------------------------------------------------------------------- */
void oracle_iso_diffusion(const oracle_params *P, int32_t iso, double *tr, double *dtracer, const double *Kfield,
                          const double *Ai_ez, const double *Ai_nz, const double *Ai_bx, const double *Ai_by,
                          const double *K_11, const double *K_22, const double *K_33, const uint8_t *maskT,
                          const uint8_t *maskW, const int32_t *kbot, const double *dxt, const double *dxu,
                          const double *dyt, const double *dyu, const double *cost, const double *cosu,
                          const double *dzt, const double *dzw, const double *int_drhodX, double *P_diss,
                          double *flux_east, double *flux_north, double *flux_top, int32_t tdma_mode) {
    const int N = P->N, M = P->M, nz = P->nz, tau = P->tau, taup1 = P->taup1;
    const size_t n3 = (size_t)N * M * nz;
    const double dt = P->dt_tracer;
    memset(flux_east, 0, n3 * 8);
    memset(flux_north, 0, n3 * 8);
    memset(flux_top, 0, n3 * 8);
    /* K1 = K_iso - K_skew, K2 = K_iso + K_skew  (:15-16, :180-188) */
#define K1(i, j, k) (iso ? (Kfield[IDX(i, j, k)] - 0.0) : (0.0 - Kfield[IDX(i, j, k)]))
#define K2(i, j, k) (iso ? (Kfield[IDX(i, j, k)] + 0.0) : (0.0 + Kfield[IDX(i, j, k)]))
    /* tr_pad (pad_z_edges): index kk in [-1, nz] clamps to [0, nz-1] */
#define TRP(i, j, kk) tr[TIDX(i, j, (kk) < 0 ? 0 : ((kk) > nz - 1 ? nz - 1 : (kk)), tau)]
#define TR(i, j, k) tr[TIDX(i, j, k, tau)]

#pragma omp parallel for schedule(static)
    for (int i = 1; i < N - 2; i++) /* east flux (:25-47) */
        for (int j = 2; j < M - 2; j++)
            for (int k = 0; k < nz; k++) {
                double diffloc = (k >= 1) ? 0.25 * (K1(i, j, k) + K1(i, j, k - 1) + K1(i + 1, j, k) + K1(i + 1, j, k - 1))
                                          : 0.5 * (K1(i, j, 0) + K1(i + 1, j, 0));
                double sumz = 0.0;
                for (int kr = 0; kr < 2; kr++)
                    for (int ip = 0; ip < 2; ip++)
                        sumz = sumz + diffloc * Ai_ez[AIDX(i, j, k, ip, kr)] *
                                          (TRP(i + ip, j, k + kr) - TRP(i + ip, j, k + kr - 1));
                flux_east[IDX(i, j, k)] =
                    sumz / (4.0 * dzt[k]) + (TR(i + 1, j, k) - TR(i, j, k)) / (cost[j] * dxu[i]) * K_11[IDX(i, j, k)];
            }
#pragma omp parallel for schedule(static)
    for (int i = 2; i < N - 2; i++) /* north flux (:52-77) */
        for (int j = 1; j < M - 2; j++)
            for (int k = 0; k < nz; k++) {
                double diffloc = (k >= 1) ? 0.25 * (K1(i, j, k) + K1(i, j, k - 1) + K1(i, j + 1, k) + K1(i, j + 1, k - 1))
                                          : 0.5 * (K1(i, j, 0) + K1(i, j + 1, 0));
                double sumz = 0.0;
                for (int kr = 0; kr < 2; kr++)
                    for (int jp = 0; jp < 2; jp++)
                        sumz = sumz + diffloc * Ai_nz[AIDX(i, j, k, jp, kr)] *
                                          (TRP(i, j + jp, k + kr) - TRP(i, j + jp, k + kr - 1));
                flux_north[IDX(i, j, k)] =
                    cosu[j] * (sumz / (4.0 * dzt[k]) + (TR(i, j + 1, k) - TR(i, j, k)) / dyu[j] * K_22[IDX(i, j, k)]);
            }
#pragma omp parallel for schedule(static)
    for (int i = 2; i < N - 2; i++) /* top flux (:85-111) */
        for (int j = 2; j < M - 2; j++)
            for (int k = 0; k < nz - 1; k++) {
                double diffloc = K2(i, j, k);
                double sumx = 0.0, sumy = 0.0;
                for (int ip = 0; ip < 2; ip++)
                    for (int kr = 0; kr < 2; kr++)
                        sumx = sumx + diffloc * Ai_bx[AIDX(i, j, k, ip, kr)] / cost[j] *
                                          (TR(i + ip, j, k + kr) - TR(i - 1 + ip, j, k + kr));
                for (int jp = 0; jp < 2; jp++)
                    for (int kr = 0; kr < 2; kr++)
                        sumy = sumy + diffloc * Ai_by[AIDX(i, j, k, jp, kr)] * cosu[j - 1 + jp] *
                                          (TR(i, j + jp, k + kr) - TR(i, j - 1 + jp, k + kr));
                flux_top[IDX(i, j, k)] = sumx / (4 * dxt[i]) + sumy / (4 * dyt[j] * cost[j]);
            }
#undef K1
#undef K2
#undef TRP
#undef TR

    /* explicit part (:116-139), dtracer += dtr, tr[taup1] += dt*dtr (:195-197) */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++)
        for (int j = 0; j < M; j++)
            for (int k = 0; k < nz; k++) {
                double mT = (double)maskT[IDX(i, j, k)];
                double e = 0.0;
                if (i >= 2 && i < N - 2 && j >= 2 && j < M - 2)
                    e = mT * ((flux_east[IDX(i, j, k)] - flux_east[IDX(i - 1, j, k)]) / (cost[j] * dxt[i]) +
                              (flux_north[IDX(i, j, k)] - flux_north[IDX(i, j - 1, k)]) / (cost[j] * dyt[j]));
                if (k == 0)
                    e += mT * flux_top[IDX(i, j, 0)] / dzt[0];
                else
                    e += mT * (flux_top[IDX(i, j, k)] - flux_top[IDX(i, j, k - 1)]) / dzt[k];
                dtracer[IDX(i, j, k)] = dtracer[IDX(i, j, k)] + e;
                if (i >= 2 && i < N - 2 && j >= 2 && j < M - 2)
                    tr[TIDX(i, j, k, taup1)] += dt * e;
            }

    /* implicit part (:142-169, :201-205) */
    if (iso) {
#pragma omp parallel
        {
        double *a = calloc(nz, 8), *b = calloc(nz, 8), *c = calloc(nz, 8), *d = calloc(nz, 8), *delta = calloc(nz, 8);
        double *be = calloc(nz, 8), *x = calloc(nz, 8), *w1 = calloc(nz, 8), *w2 = calloc(nz, 8);
        uint8_t *water = calloc(nz, 1), *edge = calloc(nz, 1);
#pragma omp for schedule(static)
        for (int i = 2; i < N - 2; i++)
            for (int j = 2; j < M - 2; j++) {
                int ks = kbot[i * M + j] - 1;
                int land = ks >= 0;
                for (int k = 0; k < nz; k++) {
                    water[k] = land && k >= ks;
                    edge[k] = land && k == ks;
                    delta[k] = (k < nz - 1) ? dt / dzw[k] * K_33[IDX(i, j, k)] : 0.0;
                    d[k] = tr[TIDX(i, j, k, taup1)];
                }
                for (int k = 0; k < nz; k++) {
                    a[k] = (k >= 1) ? -delta[k - 1] / dzt[k] : 0.0;
                    if (k >= 1 && k < nz - 1) b[k] = 1 + (delta[k] + delta[k - 1]) / dzt[k];
                    else if (k == nz - 1 && k >= 1) b[k] = 1 + delta[k - 1] / dzt[k];
                    else b[k] = 0.0;
                    be[k] = 1 + (delta[k] / dzt[k]);
                    c[k] = (k < nz - 1) ? -delta[k] / dzt[k] : 0.0;
                    if (edge[k]) b[k] = be[k]; /* solve_implicit: where(edge_mask, b_edge, b) */
                }
                solve_column(nz, a, b, c, d, water, edge, x, w1, w2, tdma_mode);
                for (int k = 0; k < nz; k++) {
                    double old = tr[TIDX(i, j, k, taup1)];
                    double nw = water[k] ? x[k] : old;
                    dtracer[IDX(i, j, k)] = dtracer[IDX(i, j, k)] + (nw - old) / dt;
                    tr[TIDX(i, j, k, taup1)] = nw;
                }
            }
        free(a); free(b); free(c); free(d); free(delta); free(be); free(x); free(w1); free(w2); free(water); free(edge);
        }
    }

    /* dissipation (:234-281; veros/core/diffusion.py:9-62) */
    if (P->enable_conserve_energy) {
        double *diss = calloc(n3, 8);
        const double fac = 0.5 * P->grav / P->rho_0;
#define X(i, j, k) int_drhodX[TIDX(i, j, k, tau)]
#pragma omp parallel for schedule(static)
        for (int i = 1; i < N - 1; i++)
            for (int j = 1; j < M - 1; j++)
                for (int k = 0; k < nz; k++)
                    diss[IDX(i, j, k)] =
                        fac * ((X(i + 1, j, k) - X(i, j, k)) * flux_east[IDX(i, j, k)] +
                               (X(i, j, k) - X(i - 1, j, k)) * flux_east[IDX(i - 1, j, k)]) / (dxt[i] * cost[j]) +
                        fac * ((X(i, j + 1, k) - X(i, j, k)) * flux_north[IDX(i, j, k)] +
                               (X(i, j, k) - X(i, j - 1, k)) * flux_north[IDX(i, j - 1, k)]) / (dyt[j] * cost[j]);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; i++)
            for (int j = 0; j < M; j++) {
                int ks = kbot[i * M + j] - 1;
                int land = ks >= 0;
                for (int k = 0; k < nz; k++) {
                    double dw;
                    if (k < nz - 1) {
                        double edge = (double)(land && k == ks), water = (double)(land && k > ks);
                        double dk = diss[IDX(i, j, k)], dk1 = diss[IDX(i, j, k + 1)];
                        dw = (0.5 * (dk + dk1) + 0.5 * (dk * dzw[k > 0 ? k - 1 : 0] / dzw[k])) * edge +
                             0.5 * (dk + dk1) * water;
                    } else {
                        dw = diss[IDX(i, j, k)] * (double)land;
                    }
                    P_diss[IDX(i, j, k)] = P_diss[IDX(i, j, k)] + dw;
                }
            }
        const double gr = -P->grav / P->rho_0;
#pragma omp parallel for schedule(static)
        for (int i = 2; i < N - 2; i++)
            for (int j = 2; j < M - 2; j++)
                for (int k = 0; k < nz - 1; k++) {
                    double fxa = (-X(i, j, k + 1) + X(i, j, k)) / dzw[k];
                    double mW = (double)maskW[IDX(i, j, k)];
                    double ft = flux_top[IDX(i, j, k)];
                    if (iso)
                        P_diss[IDX(i, j, k)] +=
                            gr * fxa * (ft * mW + K_33[IDX(i, j, k)] *
                                                      (tr[TIDX(i, j, k + 1, taup1)] - tr[TIDX(i, j, k, taup1)]) / dzw[k] * mW);
                    else
                        P_diss[IDX(i, j, k)] += gr * fxa * ft * mW;
                }
#undef X
        free(diss);
    }
}

/* ------------------------------------------------------------------------------------------
 * vertmix_tempsalt (veros/core/thermodynamics.py:248-300): implicit vertical mixing of T and S with
 * kappaH and the surface fluxes, tendencies dtemp_vmix / dsalt_vmix, then enforce_boundaries
 * (utilities.py:8-20) on temp/salt[..., taup1].  temp, salt: (N,M,nz,3); kappaH: (N,M,nz);
 * forc_*_surface: (N,M); dtemp_vmix, dsalt_vmix: (N,M,nz) outputs (every element written).
 * ---------------------------------------------------------------------------------------- */
void oracle_vertmix_tempsalt(int32_t N, int32_t M, int32_t nz, int32_t taup1, int32_t cyclic_x, double dt,
                             double *temp, double *salt, const double *kappaH, const double *forc_temp,
                             const double *forc_salt, const int32_t *kbot, const double *dzt, const double *dzw,
                             double *dtemp_vmix, double *dsalt_vmix, int32_t tdma_mode) {
    const size_t n3 = (size_t)N * M * nz;
    double *trs[2] = {temp, salt};
    double *dtrs[2] = {dtemp_vmix, dsalt_vmix};
    const double *forc[2] = {forc_temp, forc_salt};
    for (int t = 0; t < 2; t++) /* :256-257 */
        for (size_t c = 0; c < n3; c++) dtrs[t][c] = trs[t][c * 3 + taup1];
#pragma omp parallel
    {
    double *a = calloc(nz, 8), *b = calloc(nz, 8), *c = calloc(nz, 8), *d = calloc(nz, 8), *delta = calloc(nz, 8);
    double *x = calloc(nz, 8), *w1 = calloc(nz, 8), *w2 = calloc(nz, 8);
    uint8_t *water = calloc(nz, 1), *edge = calloc(nz, 1);
#pragma omp for schedule(static)
    for (int i = 2; i < N - 2; i++)
        for (int j = 2; j < M - 2; j++) {
            int ks = kbot[i * M + j] - 1;
            int land = ks >= 0;
            for (int k = 0; k < nz; k++) {
                water[k] = land && k >= ks;
                edge[k] = land && k == ks;
                delta[k] = (k < nz - 1) ? dt / dzw[k] * kappaH[IDX(i, j, k)] : 0.0; /* :267-270 */
            }
            for (int k = 0; k < nz; k++) {
                a[k] = (k >= 1) ? -delta[k - 1] / dzt[k] : 0.0;                     /* :271 */
                b[k] = (k >= 1) ? 1 + (delta[k] + delta[k - 1]) / dzt[k] : 0.0;     /* :272 */
                c[k] = (k < nz - 1) ? -delta[k] / dzt[k] : 0.0;                     /* :274 */
                if (edge[k]) b[k] = 1 + delta[k] / dzt[k];                          /* :273, b_edge */
            }
            for (int t = 0; t < 2; t++) {
                double *tr = trs[t];
                for (int k = 0; k < nz; k++) d[k] = tr[TIDX(i, j, k, taup1)];
                d[nz - 1] = d[nz - 1] + dt * forc[t][i * M + j] / dzt[nz - 1];      /* :276, :282 */
                solve_column(nz, a, b, c, d, water, edge, x, w1, w2, tdma_mode);
                for (int k = 0; k < nz; k++)
                    if (water[k]) tr[TIDX(i, j, k, taup1)] = x[k];                  /* :279, :285 */
            }
        }
    free(a); free(b); free(c); free(d); free(delta); free(x); free(w1); free(w2); free(water); free(edge);
    }
    for (int t = 0; t < 2; t++) /* :287-288 */
        for (size_t c = 0; c < n3; c++) dtrs[t][c] = (trs[t][c * 3 + taup1] - dtrs[t][c]) / dt;
    if (cyclic_x) /* enforce_boundaries: [-2:] <- [2:4], [:2] <- [-4:-2] */
        for (int t = 0; t < 2; t++)
            for (int g = 0; g < 2; g++)
                for (int j = 0; j < M; j++)
                    for (int k = 0; k < nz; k++) {
                        trs[t][TIDX(N - 2 + g, j, k, taup1)] = trs[t][TIDX(2 + g, j, k, taup1)];
                        trs[t][TIDX(g, j, k, taup1)] = trs[t][TIDX(N - 4 + g, j, k, taup1)];
                    }
}

/* ------------------------------------------------------------------------------------------
 * implicit_vert_friction (veros/core/friction.py:92-205): implicit vertical friction of u and v with
 * kappaM (two solve_implicit calls with b_edge), the du_mix / dv_mix tendencies and the dissipation
 * added to K_diss_v after ugrid_to_tgrid / vgrid_to_tgrid (veros/core/numerics.py:313-336).
 * the first extra solve_implicit caller.  Golden vectors
 * tests/golden/fric_*.npz (bit-for-bit).
 * u, v: (N,M,nz,3); kappaM, du_mix, dv_mix, K_diss_v: (N,M,nz); maskU, maskV: u8 (N,M,nz);
 * kbot: (N,M); dxt, dxu: (N); area_v, area_t: (N,M).
 * ---------------------------------------------------------------------------------------- */
void oracle_implicit_vert_friction(int32_t N, int32_t M, int32_t nz, int32_t tau, int32_t taup1, double dt_mom,
                                   double *u, double *v, const double *kappaM, const uint8_t *maskU,
                                   const uint8_t *maskV, const int32_t *kbot, const double *dzt, const double *dzw,
                                   const double *dxt, const double *dxu, const double *area_v, const double *area_t,
                                   double *du_mix, double *dv_mix, double *K_diss_v, int32_t tdma_mode) {
    const size_t n3 = (size_t)N * M * nz;
    double *vel[2] = {u, v};
    double *dmix[2] = {du_mix, dv_mix};
    const uint8_t *mask[2] = {maskU, maskV};
    for (int comp = 0; comp < 2; comp++) {
        double *diss = calloc(n3, 8); /* :100, zeros outside [1:-2, 1:-2, :-1] */
        double *w = vel[comp];
        const uint8_t *m = mask[comp];
#pragma omp parallel
        {
        double *a = calloc(nz, 8), *b = calloc(nz, 8), *c = calloc(nz, 8), *d = calloc(nz, 8), *delta = calloc(nz, 8);
        double *be = calloc(nz, 8), *fxa = calloc(nz, 8), *x = calloc(nz, 8), *w1 = calloc(nz, 8), *w2 = calloc(nz, 8);
        uint8_t *water = calloc(nz, 1), *edge = calloc(nz, 1);
#pragma omp for schedule(static)
        for (int i = 1; i < N - 2; i++)
            for (int j = 1; j < M - 2; j++) {
                const int in = comp == 0 ? i + 1 : i, jn = comp == 0 ? j : j + 1; /* the cell on the other side of the face */
                const int kb0 = kbot[i * M + j], kb1 = kbot[in * M + jn];
                const int ks = (kb0 > kb1 ? kb0 : kb1) - 1; /* :111 / :158, create_water_masks */
                const int land = ks >= 0;
                for (int k = 0; k < nz; k++) {
                    water[k] = land && k >= ks;
                    edge[k] = land && k == ks;
                    if (k < nz - 1) {
                        fxa[k] = 0.5 * (kappaM[IDX(i, j, k)] + kappaM[IDX(in, jn, k)]);                           /* :114 */
                        delta[k] = dt_mom / dzw[k] * fxa[k] * (double)m[IDX(i, j, k + 1)] * (double)m[IDX(i, j, k)]; /* :115-117 */
                    } else {
                        fxa[k] = 0.0;
                        delta[k] = 0.0;
                    }
                }
                for (int k = 0; k < nz; k++) {
                    a[k] = (k >= 1) ? -delta[k - 1] / dzt[k] : 0.0;      /* :118 */
                    b[k] = (k >= 1) ? 1 + delta[k - 1] / dzt[k] : 0.0;   /* :119 */
                    if (k >= 1 && k < nz - 1) b[k] = b[k] + delta[k] / dzt[k]; /* :120 */
                    be[k] = 1 + delta[k] / dzt[k];                       /* :121 */
                    c[k] = -delta[k] / dzt[k];                           /* :122 (u) / :173-174 (v: last level +0) */
                    d[k] = w[TIDX(i, j, k, tau)];                        /* :123 */
                    if (edge[k]) b[k] = be[k];
                }
                if (comp == 1) c[nz - 1] = 0.0;
                solve_column(nz, a, b, c, d, water, edge, x, w1, w2, tdma_mode); /* :125 */
                for (int k = 0; k < nz; k++) {
                    if (water[k]) w[TIDX(i, j, k, taup1)] = x[k];                                     /* :126 */
                    dmix[comp][IDX(i, j, k)] = (w[TIDX(i, j, k, taup1)] - w[TIDX(i, j, k, tau)]) / dt_mom; /* :127-129 */
                }
                for (int k = 0; k < nz - 1; k++) { /* :134-148 */
                    const double ft = fxa[k] * (w[TIDX(i, j, k + 1, taup1)] - w[TIDX(i, j, k, taup1)]) / dzw[k] *
                                      (double)m[IDX(i, j, k + 1)] * (double)m[IDX(i, j, k)];
                    diss[IDX(i, j, k)] = (w[TIDX(i, j, k + 1, tau)] - w[TIDX(i, j, k, tau)]) * ft / dzw[k];
                }
            }
        free(a); free(b); free(c); free(d); free(delta); free(be); free(fxa); free(x); free(w1); free(w2);
        free(water); free(edge);
        }
        /* ugrid_to_tgrid / vgrid_to_tgrid on the whole array, then K_diss_v += (:150-151, :202-203) */
        for (int i = 0; i < N; i++)
            for (int j = 0; j < M; j++)
                for (int k = 0; k < nz; k++) {
                    double t = diss[IDX(i, j, k)];
                    if (comp == 0 && i >= 2 && i < N - 2)
                        t = (dxu[i] * diss[IDX(i, j, k)] + dxu[i - 1] * diss[IDX(i - 1, j, k)]) / (2 * dxt[i]);
                    if (comp == 1 && j >= 2 && j < M - 2)
                        t = (area_v[i * M + j] * diss[IDX(i, j, k)] + area_v[i * M + j - 1] * diss[IDX(i, j - 1, k)]) /
                            (2 * area_t[i * M + j]);
                    K_diss_v[IDX(i, j, k)] = K_diss_v[IDX(i, j, k)] + t;
                }
        free(diss);
    }
}

/* ------------------------------------------------------------------------------------------
 * isoneutral_diag_streamfunction_kernel (veros/core/isoneutral/isoneutral.py:232-258): B1_gm, B2_gm from
 * K_gm and the Ai_ez / Ai_nz the path just produced (consumer).
 * np.sum(..., axis=(3, 4)) of a C-contiguous (...,2,2) block adds its four elements in memory order.
 * ---------------------------------------------------------------------------------------- */
void oracle_diag_streamfunction(int32_t N, int32_t M, int32_t nz, const double *K_gm, const double *Ai_ez,
                                const double *Ai_nz, double *B1_gm, double *B2_gm) {
    for (int i = 1; i < N - 2; i++)
        for (int j = 1; j < M - 2; j++)
            for (int k = 0; k < nz; k++) {
                const int km = k > 0 ? k - 1 : 0; /* pad_z_edges */
                if (j >= 2) { /* :241-245, [1:-2, 2:-2] */
                    const double diffloc = 0.25 * (K_gm[IDX(i, j, k)] + K_gm[IDX(i, j, km)] + K_gm[IDX(i + 1, j, k)] +
                                                   K_gm[IDX(i + 1, j, km)]);
                    const double *A = Ai_ez + IDX(i, j, k) * 4;
                    const double s = ((A[0] + A[1]) + A[2]) + A[3];
                    B2_gm[IDX(i, j, k)] = 0.25 * diffloc * s;
                }
                if (i >= 2) { /* :250-254, [2:-2, 1:-2] */
                    const double diffloc = 0.25 * (K_gm[IDX(i, j, k)] + K_gm[IDX(i, j, km)] + K_gm[IDX(i, j + 1, k)] +
                                                   K_gm[IDX(i, j + 1, km)]);
                    const double *A = Ai_nz + IDX(i, j, k) * 4;
                    const double s = ((A[0] + A[1]) + A[2]) + A[3];
                    B1_gm[IDX(i, j, k)] = -0.25 * diffloc * s;
                }
            }
}

/* ------------------------------------------------------------------------------------------
 * set_eke_diffusivities_kernel (veros/core/eke.py:34-85): the producer of K_gm and K_iso
 * (SURVEY.md 8(f) rank 4).  The column sum of :44-51 is NumPy's add.reduce over a contiguous axis: the
 * accumulator starts from the identity 0 and the run goes through pairwise summation (8 interleaved
 * accumulators up to 128 elements, recursive halving above; numpy/_core/src/umath/loops_utils.h.src,
 * numpy 2.3 as installed) -- sum_variant 1, pinned by tests/golden/eke_*.npz (variant 0, first element
 * outside the pairwise run, does NOT reproduce the reference).
 * ---------------------------------------------------------------------------------------- */
static double np_pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int q = 0; q < 8; q++) r[q] += a[i + q];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

double oracle_np_sum(const double *a, int64_t n, int32_t variant) {
    if (n == 0) return 0.0;
    if (variant == 0) return a[0] + np_pairwise_sum(a + 1, n - 1);
    return 0.0 + np_pairwise_sum(a, n);
}

void oracle_set_eke_diffusivities(int32_t N, int32_t M, int32_t nz, int32_t tau, int32_t enable_eke,
                                  int32_t enable_eke_isopycnal_diffusion, double pi, double eke_lmin, double eke_cross,
                                  double eke_crhin, double eke_k_max, double eke_c_k, double K_gm_0, double K_iso_0,
                                  const double *Nsqr, const double *eke, const uint8_t *maskW, const double *dzw,
                                  const double *coriolis_t, const double *beta, double *L_rossby, double *L_rhines,
                                  double *eke_len, double *sqrteke, double *K_gm, double *K_iso, int32_t sum_variant) {
    const size_t n3 = (size_t)N * M * nz;
    if (!enable_eke) { /* :73-82 */
        for (size_t c = 0; c < n3; c++) {
            K_gm[c] = K_gm_0;
            K_iso[c] = K_iso_0;
        }
        return;
    }
    double *term = malloc(8 * (size_t)nz);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < M; j++) {
            for (int k = 0; k < nz; k++) { /* :44-50 */
                const double n2 = Nsqr[TIDX(i, j, k, tau)];
                term[k] = sqrt(n2 > 0.0 ? n2 : 0.0) * dzw[k] * (double)maskW[IDX(i, j, k)] / pi;
            }
            const double C = oracle_np_sum(term, nz, sum_variant);
            const double f = fabs(coriolis_t[i * M + j]);
            const double b2 = 2 * beta[i * M + j];
            const double l1 = C / (f > 1e-16 ? f : 1e-16);
            const double l2 = sqrt(C / (b2 > 1e-16 ? b2 : 1e-16));
            const double Lr = l1 < l2 ? l1 : l2; /* :52-54 */
            L_rossby[i * M + j] = Lr;
            const double bt = beta[i * M + j] > 1e-16 ? beta[i * M + j] : 1e-16;
            for (int k = 0; k < nz; k++) {
                const double e = eke[TIDX(i, j, k, tau)];
                const double se = sqrt(e > 0.0 ? e : 0.0); /* :59 */
                const double lrh = sqrt(se / bt);           /* :60 */
                const double x1 = eke_cross * Lr, x2 = eke_crhin * lrh;
                const double mn = x1 < x2 ? x1 : x2;
                const double len = eke_lmin > mn ? eke_lmin : mn; /* :61-64 */
                const double kg = eke_c_k * len * se;
                sqrteke[IDX(i, j, k)] = se;
                L_rhines[IDX(i, j, k)] = lrh;
                eke_len[IDX(i, j, k)] = len;
                K_gm[IDX(i, j, k)] = eke_k_max < kg ? eke_k_max : kg; /* :65 */
                K_iso[IDX(i, j, k)] = enable_eke_isopycnal_diffusion ? K_gm[IDX(i, j, k)] : K_iso_0; /* :75-78 */
            }
        }
    free(term);
}

/* ------------------------------------------------------------------------------------------
 * advect_tracer (veros/core/thermodynamics.py:10-40) with adv_flux_2nd / adv_flux_superbee
 * (veros/core/advection.py:8-115), advect_temperature / advect_salinity (:43-62) and the Adams-Bashforth
 * step of :223-245 -- the producer of temp/salt[..., taup1] right before the isoneutral path
 * (SURVEY.md 8(f) rank 4).  Golden vectors tests/golden/adv_*.npz (bit-for-bit).
 * ---------------------------------------------------------------------------------------- */
static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* one superbee face flux: values at sm1, s, sp1, sp2 and masks at sm1, s, sp1 (advection.py:39-49) */
static double superbee_face(double vm1, double v0, double v1, double v2, double mm1, double m0, double m1, double vel,
                            double velc, double dt, double dx) {
    const double eps = 1e-20;
    const double rjp = (v2 - v1) * m1, rj = (v1 - v0) * m0, rjm = (v0 - vm1) * mm1;
    const double num = vel > 0.0 ? rjm : rjp;
    const double den = fabs(rj) < eps ? eps : rj;
    const double c = num / den;
    const double a1 = clipd(2 * c, 0, 1), a2 = clipd(c, 0, 2);
    const double cr = a1 > a2 ? a1 : a2;
    const double uCFL = fabs(velc * dt / dx);
    return velc * (v1 + v0) * 0.5 - fabs(velc) * ((1.0 - cr) + uCFL * cr) * rj * 0.5;
}

/* tr: (N,M,nz,3) at time level `lev`; u, v, w at time level tau; dtr: (N,M,nz) fully written */
void oracle_advect_tracer(int32_t N, int32_t M, int32_t nz, int32_t superbee, int32_t tau, int32_t lev, double dt_tracer,
                          const double *tr, const double *u, const double *v, const double *w, const uint8_t *maskT,
                          const uint8_t *maskU, const uint8_t *maskV, const uint8_t *maskW, const double *dxt,
                          const double *dyt, const double *dzt, const double *cost, const double *cosu, double *dtr) {
    const size_t n3 = (size_t)N * M * nz;
    double *fe = calloc(n3, 8), *fn = calloc(n3, 8), *ft = calloc(n3, 8);
#define VAR(i, j, k) tr[TIDX(i, j, k, lev)]
#define KC(k) ((k) < 0 ? 0 : ((k) > nz - 1 ? nz - 1 : (k)))
#pragma omp parallel for schedule(static)
    for (int i = 1; i < N - 2; i++)
        for (int j = 1; j < M - 2; j++)
            for (int k = 0; k < nz; k++) {
                if (j >= 2) { /* east face, [1:-2, 2:-2, :] */
                    const double vel = u[TIDX(i, j, k, tau)];
                    if (superbee)
                        fe[IDX(i, j, k)] = superbee_face(VAR(i - 1, j, k), VAR(i, j, k), VAR(i + 1, j, k), VAR(i + 2, j, k),
                                                         maskU[IDX(i - 1, j, k)], maskU[IDX(i, j, k)], maskU[IDX(i + 1, j, k)],
                                                         vel, vel, dt_tracer, cost[j] * dxt[i]);
                    else
                        fe[IDX(i, j, k)] = 0.5 * (VAR(i, j, k) + VAR(i + 1, j, k)) * vel * maskU[IDX(i, j, k)];
                }
                if (i >= 2) { /* north face, [2:-2, 1:-2, :] */
                    const double vel = v[TIDX(i, j, k, tau)];
                    if (superbee)
                        fn[IDX(i, j, k)] = superbee_face(VAR(i, j - 1, k), VAR(i, j, k), VAR(i, j + 1, k), VAR(i, j + 2, k),
                                                         maskV[IDX(i, j - 1, k)], maskV[IDX(i, j, k)], maskV[IDX(i, j + 1, k)],
                                                         vel, vel * cosu[j], dt_tracer, cost[j] * dyt[j]);
                    else
                        fn[IDX(i, j, k)] = cosu[j] * 0.5 * (VAR(i, j, k) + VAR(i, j + 1, k)) * vel * maskV[IDX(i, j, k)];
                }
                if (i >= 2 && j >= 2 && k < nz - 1) { /* top face, [2:-2, 2:-2, :-1]; pad_z_edges clamps */
                    const double vel = w[TIDX(i, j, k, tau)];
                    if (superbee)
                        ft[IDX(i, j, k)] = superbee_face(VAR(i, j, KC(k - 1)), VAR(i, j, k), VAR(i, j, k + 1), VAR(i, j, KC(k + 2)),
                                                         maskW[IDX(i, j, KC(k - 1))], maskW[IDX(i, j, k)], maskW[IDX(i, j, k + 1)],
                                                         vel, vel, dt_tracer, dzt[k]);
                    else
                        ft[IDX(i, j, k)] = 0.5 * (VAR(i, j, k) + VAR(i, j, k + 1)) * vel * maskW[IDX(i, j, k)];
                }
            }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++)
        for (int j = 0; j < M; j++)
            for (int k = 0; k < nz; k++) {
                double d = 0.0;
                if (i >= 2 && i < N - 2 && j >= 2 && j < M - 2) /* thermodynamics.py:24-35 */
                    d = maskT[IDX(i, j, k)] * (-(fe[IDX(i, j, k)] - fe[IDX(i - 1, j, k)]) / (cost[j] * dxt[i]) -
                                               (fn[IDX(i, j, k)] - fn[IDX(i, j - 1, k)]) / (cost[j] * dyt[j]));
                const double mt = -1 * (int)maskT[IDX(i, j, k)];
                if (k == 0)
                    d = d + mt * ft[IDX(i, j, 0)] / dzt[0]; /* :36 */
                else
                    d = d + mt * (ft[IDX(i, j, k)] - ft[IDX(i, j, k - 1)]) / dzt[k]; /* :37-39 */
                dtr[IDX(i, j, k)] = d;
            }
#undef VAR
#undef KC
    free(fe); free(fn); free(ft);
}

/* Adams-Bashforth step (thermodynamics.py:223-245) of one tracer over the whole array */
void oracle_adams_bashforth(int32_t N, int32_t M, int32_t nz, int32_t tau, int32_t taup1, int32_t taum1, double dt_tracer,
                            double AB_eps, double *tr, const double *dtr, const uint8_t *maskT) {
    const size_t n3 = (size_t)N * M * nz;
    for (size_t c = 0; c < n3; c++)
        tr[c * 3 + taup1] = tr[c * 3 + tau] +
                            dt_tracer * ((1.5 + AB_eps) * dtr[c * 3 + tau] - (0.5 + AB_eps) * dtr[c * 3 + taum1]) * maskT[c];
}
