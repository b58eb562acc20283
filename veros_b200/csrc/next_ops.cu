// Neighbours of the isoneutral path:
//
//  * implicit_vert_friction (veros/core/friction.py:92-205) -- the first other caller of solve_implicit, with the
//    coefficient assembly fused into the solve: per velocity component one kernel on tiles of whole columns
//      phase B  per cell: delta from kappaM and the U / V masks, matrix rows (a_tri, b_tri, b_tri_edge, c_tri), rhs
//      phase C  per column: LAPACK dgtsv replay (tdma_device.cuh)
//      phase D  per cell: where(water, sol, vel), d{u,v}_mix, the raw dissipation of friction.py:134-148
//    and one element-wise kernel for ugrid_to_tgrid / vgrid_to_tgrid (numerics.py:313-336) and K_diss_v +=.
//    The reference materialises a_tri, b_tri, b_tri_edge, c_tri, d_tri, delta, flux_top, diss and two masks per
//    component and calls the generic solve; here none of them exists in HBM.
//  * isoneutral_diag_streamfunction_kernel (veros/core/isoneutral/isoneutral.py:232-258) -- consumer of Ai_ez / Ai_nz.
//  * set_eke_diffusivities_kernel (veros/core/eke.py:34-85) -- producer of K_gm / K_iso, including NumPy's pairwise
//    column sum order.
// Every operation is an explicitly rounded intrinsic in the reference's order: bit-identical to the NumPy backend.
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

constexpr int kFricBlock = 128;

struct FricArgs {
    int N, M, nz;
    double* vel[2];          // u, v (N,M,nz,3) in/out (taup1 level written)
    double* dmix[2];         // du_mix, dv_mix (N,M,nz) in/out
    const int32_t *tau, *taup1;
    const double* kappaM;
    const uint8_t* mask[2];  // maskU, maskV
    const int32_t* kbot;
    const double *dzt, *dzw;
    double* diss[2];         // scratch: raw dissipation of each component (N,M,nz)
    double dt_mom;
};

// COMP 0: zonal momentum (neighbour cell i+1), COMP 1: meridional momentum (neighbour cell j+1)
template <int COMP>
__global__ void __launch_bounds__(kFricBlock, 6)
friction_kernel(const FricArgs a, const int cols, const int pitch) {
    extern __shared__ double sm[];
    const int N = a.N, M = a.M, nz = a.nz;
    const int i = 1 + blockIdx.y;                 // planes [1, N-2)
    const int j0 = 1 + blockIdx.x * cols;         // columns [1, M-2)
    const int ncols = min(cols, (M - 2) - j0);
    const int ncells = ncols * nz;
    const int tau = *a.tau, taup1 = *a.taup1;
    const double dt = a.dt_mom;
    const Divisor ddt = make_divisor(dt);
    const size_t base = ((size_t)i * M + j0) * nz;
    const size_t nb = COMP == 0 ? (size_t)M * nz : (size_t)nz;  // offset of the cell across the face
    double* __restrict__ vel = a.vel[COMP];
    const uint8_t* __restrict__ mk = a.mask[COMP];

    Divisor* ddzt = reinterpret_cast<Divisor*>(sm);  // dzt[k]
    Divisor* ddzw = ddzt + nz;                        // dzw[k]
    double* dt_dzw = reinterpret_cast<double*>(ddzw + nz);  // dt_mom / dzw[k]
    const int tile = cols * pitch;
    double* L = dt_dzw + nz;
    double* D = L + tile;
    double* U = D + tile;
    double* R = U + tile;
    double* W1 = R + tile;  // vel[..., taup1] before the solve
    int* ksv = reinterpret_cast<int*>(W1 + tile);

    for (int k = threadIdx.x; k < nz; k += kFricBlock) {
        ddzt[k] = make_divisor(a.dzt[k]);
        ddzw[k] = make_divisor(a.dzw[k]);
        dt_dzw[k] = strict::div(dt, a.dzw[k]);
    }
    for (int q = threadIdx.x; q < ncols; q += kFricBlock) {
        const int j = j0 + q;
        const int kb0 = a.kbot[i * M + j], kb1 = COMP == 0 ? a.kbot[(i + 1) * M + j] : a.kbot[i * M + j + 1];
        ksv[q] = max(kb0, kb1) - 1;  // friction.py:111 / :158 -> create_water_masks
    }
    __syncthreads();

    // delta[k] of friction.py:115-117 (0 at the top level)
    auto delta_at = [&](size_t c, int k) {
        if (k >= nz - 1) return 0.0;
        const double fxa = mul(0.5, add(__ldg(a.kappaM + c), __ldg(a.kappaM + c + nb)));
        return mul(mul(mul(dt_dzw[k], fxa), (double)mk[c + 1]), (double)mk[c]);
    };

    // ---- phase B -------------------------------------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < ncells; idx += kFricBlock) {
        const int q = idx / nz, k = idx - q * nz;
        const size_t c = base + idx;
        const int s = q * pitch + k;
        const int ks = ksv[q];
        const double del = delta_at(c, k);
        const double delm = k > 0 ? delta_at(c - 1, k - 1) : 0.0;
        double diag;
        if (k == ks)
            diag = add(1.0, strict::div(del, ddzt[k]));                     // b_tri_edge, :121
        else {
            diag = add(1.0, strict::div(delm, ddzt[k]));                    // :119
            if (k < nz - 1) diag = add(diag, strict::div(del, ddzt[k]));    // :120
        }
        D[s] = diag;
        U[s] = (k < nz - 1) ? strict::div(-del, ddzt[k]) : 0.0;             // c_tri, :122
        if (k > 0) L[s - 1] = (k > ks) ? strict::div(-delm, ddzt[k]) : 0.0;  // a_tri, :118; 0 on the edge row
        R[s] = vel[c * 3 + tau];                                             // d_tri, :123
        W1[s] = vel[c * 3 + taup1];
    }
    __syncthreads();

    // ---- phase C -------------------------------------------------------------------------------------------------
    for (int q = threadIdx.x; q < ncols; q += kFricBlock) {
        const int ks = ksv[q];
        if (ks >= 0) {
            const int o = q * pitch;
            dgtsv_column<1>(ks, nz, L + o, D + o, U + o, R + o, nullptr);
        }
    }
    __syncthreads();

    // ---- phase D -------------------------------------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < ncells; idx += kFricBlock) {
        const int q = idx / nz, k = idx - q * nz;
        const size_t c = base + idx;
        const int s = q * pitch + k;
        const int ks = ksv[q];
        const bool land = ks >= 0;
        const bool water = land && k >= ks;
        const double v0 = vel[c * 3 + tau];
        const double v1 = water ? R[s] : W1[s];  // :126
        if (water) vel[c * 3 + taup1] = v1;
        a.dmix[COMP][c] = strict::div(sub(v1, v0), ddt);  // :127-129
        double ds = 0.0;
        if (k < nz - 1) {  // :134-148
            const bool water_up = land && k + 1 >= ks;
            const double v1u = water_up ? R[s + 1] : W1[s + 1];
            const double v0u = vel[(c + 1) * 3 + tau];
            const double fxa = mul(0.5, add(__ldg(a.kappaM + c), __ldg(a.kappaM + c + nb)));
            const double ft = mul(mul(strict::div(mul(fxa, sub(v1u, v1)), ddzw[k]), (double)mk[c + 1]), (double)mk[c]);
            ds = strict::div(mul(sub(v0u, v0), ft), ddzw[k]);
        }
        a.diss[COMP][c] = ds;
    }
}

// ugrid_to_tgrid / vgrid_to_tgrid of the two raw dissipation fields (zero outside [1:-2, 1:-2, :-1]) and K_diss_v +=
__global__ void __launch_bounds__(256)
friction_finish_kernel(int N, int M, int nz, const double* __restrict__ du, const double* __restrict__ dv,
                       const double* __restrict__ dxt, const double* __restrict__ dxu, const double* __restrict__ area_v,
                       const double* __restrict__ area_t, double* __restrict__ K_diss_v) {
    const size_t n3 = (size_t)N * M * nz;
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n3) return;
    const size_t plane = (size_t)M * nz;
    const int i = (int)(c / plane);
    const int j = (int)((c - (size_t)i * plane) / nz);
    auto raw = [&](const double* d, int ii, int jj, size_t cc) {
        return (ii >= 1 && ii < N - 2 && jj >= 1 && jj < M - 2) ? d[cc] : 0.0;
    };
    double tu = raw(du, i, j, c);
    if (i >= 2 && i < N - 2)
        tu = __ddiv_rn(add(mul(dxu[i], tu), mul(dxu[i - 1], raw(du, i - 1, j, c - plane))), mul(2.0, dxt[i]));
    double tv = raw(dv, i, j, c);
    if (j >= 2 && j < M - 2)
        tv = __ddiv_rn(add(mul(area_v[i * M + j], tv), mul(area_v[i * M + j - 1], raw(dv, i, j - 1, c - nz))),
                       mul(2.0, area_t[i * M + j]));
    K_diss_v[c] = add(add(K_diss_v[c], tu), tv);
}

// ---- isoneutral_diag_streamfunction_kernel ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
streamfunction_kernel(int N, int M, int nz, const double* __restrict__ K_gm, const double* __restrict__ Ai_ez,
                      const double* __restrict__ Ai_nz, double* __restrict__ B1_gm, double* __restrict__ B2_gm) {
    const size_t n3 = (size_t)N * M * nz;
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n3) return;
    const size_t plane = (size_t)M * nz;
    const int i = (int)(c / plane);
    const int j = (int)((c - (size_t)i * plane) / nz);
    const int k = (int)(c % nz);
    if (i < 1 || i >= N - 2 || j < 1 || j >= M - 2) return;
    const size_t cm = k > 0 ? c - 1 : c;  // pad_z_edges
    auto sum4 = [](const double* A) {  // np.sum over the trailing (2, 2) block: memory order
        const double2 x = __ldg(reinterpret_cast<const double2*>(A)), y = __ldg(reinterpret_cast<const double2*>(A) + 1);
        return add(add(add(x.x, x.y), y.x), y.y);
    };
    if (j >= 2) {
        const double diffloc = mul(0.25, add(add(add(K_gm[c], K_gm[cm]), K_gm[c + plane]), K_gm[cm + plane]));
        B2_gm[c] = mul(mul(0.25, diffloc), sum4(Ai_ez + c * 4));
    }
    if (i >= 2) {
        const double diffloc = mul(0.25, add(add(add(K_gm[c], K_gm[cm]), K_gm[c + nz]), K_gm[cm + nz]));
        B1_gm[c] = mul(mul(-0.25, diffloc), sum4(Ai_nz + c * 4));
    }
}

// ---- set_eke_diffusivities_kernel ----------------------------------------------------------------------------------
struct EkeArgs {
    int N, M, nz;
    const double *Nsqr, *eke;  // (N,M,nz,3)
    const int32_t* tau;
    const uint8_t* maskW;
    const double *dzw, *coriolis_t, *beta;
    double *L_rossby, *L_rhines, *eke_len, *sqrteke, *K_gm, *K_iso;
    double pi, eke_lmin, eke_cross, eke_crhin, eke_k_max, eke_c_k, K_gm_0, K_iso_0;
    int enable_eke, iso_from_gm;
};

// NumPy's pairwise summation of a contiguous run (numpy/_core/src/umath/loops_utils.h.src): sequential below 8
// elements, 8 interleaved accumulators up to 128, recursive halving (multiples of 8) above.
__device__ double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int q = 0; q < n; ++q) res = add(res, a[q]);
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = a[q];
        int m = 8;
        for (; m < n - (n % 8); m += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = add(r[q], a[m + q]);
        }
        double res = add(add(add(r[0], r[1]), add(r[2], r[3])), add(add(r[4], r[5]), add(r[6], r[7])));
        for (; m < n; ++m) res = add(res, a[m]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return add(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// one CTA per tile of columns: terms of the Rossby-radius sum staged in shared memory with coalesced loads, one thread
// per column adds them in NumPy's order, then every cell of the tile gets its diffusivities
__global__ void __launch_bounds__(128)
eke_kernel(const EkeArgs a, const int cols, const int pitch) {
    extern __shared__ double sm[];
    const int nz = a.nz;
    const size_t ncol_total = (size_t)a.N * a.M;
    const size_t col0 = (size_t)blockIdx.x * cols;
    const int ncols = (int)min((size_t)cols, ncol_total - col0);
    const int ncells = ncols * nz;
    const size_t base = col0 * nz;
    double* term = sm;
    double* Lr = term + (size_t)cols * pitch;
    if (!a.enable_eke) {  // eke.py:73-82
        for (int idx = threadIdx.x; idx < ncells; idx += blockDim.x) {
            a.K_gm[base + idx] = a.K_gm_0;
            a.K_iso[base + idx] = a.K_iso_0;
        }
        return;
    }
    const int tau = *a.tau;
    for (int idx = threadIdx.x; idx < ncells; idx += blockDim.x) {
        const int q = idx / nz, k = idx - q * nz;
        const size_t c = base + idx;
        const double n2 = a.Nsqr[c * 3 + tau];
        const double t = mul(mul(sqrt(fmax(0.0, n2)), a.dzw[k]), (double)a.maskW[c]);
        term[q * pitch + k] = __ddiv_rn(t, a.pi);  // eke.py:44-50
    }
    __syncthreads();
    for (int q = threadIdx.x; q < ncols; q += blockDim.x) {
        const double C = add(0.0, np_pairwise_sum(term + q * pitch, nz));
        const size_t c2 = col0 + q;
        const double l1 = __ddiv_rn(C, fmax(fabs(a.coriolis_t[c2]), 1e-16));
        const double l2 = sqrt(__ddiv_rn(C, fmax(mul(2.0, a.beta[c2]), 1e-16)));
        const double L = fmin(l1, l2);  // eke.py:52-54
        a.L_rossby[c2] = L;
        Lr[q] = L;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < ncells; idx += blockDim.x) {
        const int q = idx / nz;
        const size_t c = base + idx;
        const double se = sqrt(fmax(0.0, a.eke[c * 3 + tau]));                             // :59
        const double lrh = sqrt(__ddiv_rn(se, fmax(a.beta[col0 + q], 1e-16)));             // :60
        const double len = fmax(a.eke_lmin, fmin(mul(a.eke_cross, Lr[q]), mul(a.eke_crhin, lrh)));  // :61-64
        const double kg = fmin(a.eke_k_max, mul(mul(a.eke_c_k, len), se));                 // :65
        a.sqrteke[c] = se;
        a.L_rhines[c] = lrh;
        a.eke_len[c] = len;
        a.K_gm[c] = kg;
        a.K_iso[c] = a.iso_from_gm ? kg : a.K_iso_0;                                       // :75-78
    }
}

}  // namespace

void launch_implicit_vert_friction(cudaStream_t s, int N, int M, int nz, double dt_mom, void** B) {
    if (N < 4 || M < 4 || nz < 1) return;
    const size_t n3 = (size_t)N * M * nz;
    FricArgs a;
    a.N = N;
    a.M = M;
    a.nz = nz;
    a.vel[0] = (double*)B[17];
    a.vel[1] = (double*)B[18];
    a.dmix[0] = (double*)B[19];
    a.dmix[1] = (double*)B[20];
    a.tau = (const int32_t*)B[5];
    a.taup1 = (const int32_t*)B[6];
    a.kappaM = (const double*)B[7];
    a.mask[0] = (const uint8_t*)B[8];
    a.mask[1] = (const uint8_t*)B[9];
    a.kbot = (const int32_t*)B[10];
    a.dzt = (const double*)B[11];
    a.dzw = (const double*)B[12];
    a.diss[0] = (double*)B[22];
    a.diss[1] = a.diss[0] + n3;
    a.dt_mom = dt_mom;
    const int pitch = (nz + 1) | 1;
    int cols = max(1, 640 / nz);
    cols = min(cols, M - 3);
    const int want_tiles = 4 * 148;
    if (((M - 3 + cols - 1) / cols) * (N - 3) < want_tiles) {
        const int per_row = (want_tiles + (N - 3) - 1) / (N - 3);
        cols = max(1, (M - 3 + per_row - 1) / per_row);
    }
    const size_t smem = (size_t)nz * (2 * sizeof(Divisor) + 8) + 8 * ((size_t)5 * cols * pitch + (cols + 1) / 2 + 1);
    dim3 grid((M - 3 + cols - 1) / cols, N - 3);
    allow_big_smem(friction_kernel<0>, 200 * 1024);
    allow_big_smem(friction_kernel<1>, 200 * 1024);
    friction_kernel<0><<<grid, kFricBlock, smem, s>>>(a, cols, pitch);
    friction_kernel<1><<<grid, kFricBlock, smem, s>>>(a, cols, pitch);
    count_launch(2);
    if (!check_launch("friction_kernel")) return;
    friction_finish_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, s>>>(
        N, M, nz, a.diss[0], a.diss[1], (const double*)B[13], (const double*)B[14], (const double*)B[15],
        (const double*)B[16], (double*)B[21]);
    count_launch();
    check_launch("friction_finish_kernel");
}

void launch_diag_streamfunction(cudaStream_t s, int N, int M, int nz, void** B) {
    if (N < 4 || M < 4 || nz < 1) return;
    const size_t n3 = (size_t)N * M * nz;
    streamfunction_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, s>>>(N, M, nz, (const double*)B[0], (const double*)B[1],
                                                                      (const double*)B[2], (double*)B[5], (double*)B[6]);
    count_launch();
    check_launch("streamfunction_kernel");
}

void launch_set_eke_diffusivities(cudaStream_t s, const VerosB200EkeDescriptor* d, void** B) {
    const int N = d->nx_tot, M = d->ny_tot, nz = d->nz;
    if (N < 1 || M < 1 || nz < 1) return;
    EkeArgs a;
    a.N = N;
    a.M = M;
    a.nz = nz;
    a.Nsqr = (const double*)B[0];
    a.eke = (const double*)B[1];
    a.tau = (const int32_t*)B[2];
    a.maskW = (const uint8_t*)B[3];
    a.dzw = (const double*)B[4];
    a.coriolis_t = (const double*)B[5];
    a.beta = (const double*)B[6];
    a.L_rossby = (double*)B[7];
    a.L_rhines = (double*)B[8];
    a.eke_len = (double*)B[9];
    a.sqrteke = (double*)B[10];
    a.K_gm = (double*)B[11];
    a.K_iso = (double*)B[12];
    a.pi = d->pi;
    a.eke_lmin = d->eke_lmin;
    a.eke_cross = d->eke_cross;
    a.eke_crhin = d->eke_crhin;
    a.eke_k_max = d->eke_k_max;
    a.eke_c_k = d->eke_c_k;
    a.K_gm_0 = d->K_gm_0;
    a.K_iso_0 = d->K_iso_0;
    a.enable_eke = d->enable_eke;
    a.iso_from_gm = d->enable_eke && d->enable_eke_isopycnal_diffusion;
    const int pitch = nz | 1;
    int cols = max(1, min(64, (40 * 1024) / (8 * pitch)));
    const size_t ncol = (size_t)N * M;
    const size_t smem = 8 * ((size_t)cols * pitch + cols);
    allow_big_smem(eke_kernel, 100 * 1024);
    eke_kernel<<<(unsigned)((ncol + cols - 1) / cols), 128, smem, s>>>(a, cols, pitch);
    count_launch();
    check_launch("eke_kernel");
}

}  // namespace vb
