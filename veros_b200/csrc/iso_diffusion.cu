// isoneutral_diffusion / isoneutral_skew_diffusion (veros/core/isoneutral/diffusion.py:9-307,
// dissipation helpers veros/core/diffusion.py:9-62).
//
// Two kernels per call:
//   flux_kernel    one thread per cell: the Griffies triad fluxes on the east / north / top faces
//                  (diffusion.py:9-113) for one tracer, written to scratch.
//   update_kernel  one CTA per tile of whole water columns: explicit flux divergence
//                  (diffusion.py:116-139), tracer update, the implicit K_33 solve
//                  (diffusion.py:142-169 + utilities.py:38-59 + operators.py:60-77) out of shared
//                  memory with one thread per column, tendency and dissipation (diffusion.py:196-281).
//                  With two tracers (temp and salt of one step) the column matrix, which depends on
//                  K_33 only, is factorised once and applied to both right-hand sides.
//
// Arithmetic is "strict": every NumPy ufunc call of the reference is one explicitly rounded
// operation here, in the reference's order, and the column solve replays dgtsv (tdma_device.cuh),
// so on identical inputs the outputs are bit-identical to the reference's NumPy backend.  Divisions
// by grid metrics use correctly rounded reciprocals tabulated once per CTA (strict.cuh).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "strict.cuh"
#include "tdma_device.cuh"

namespace vb {

using strict::add;
using strict::Divisor;
using strict::make_divisor;
using strict::mul;
using strict::sub;

namespace {

__device__ __forceinline__ double ldt(const double* tr, size_t cell) { return __ldg(tr + cell * 3); }

// ------------------------------------------------------------------------------------------------
// flux_kernel: grid (ceil(M*nz / blockDim), N); block b of plane i handles flattened cells
// p in [b*blockDim, (b+1)*blockDim) of that plane (contiguous in memory).
// Dynamic shared memory: per-level and per-row metric tables (divisor + reciprocal).
// ------------------------------------------------------------------------------------------------
template <bool SKEW>
__global__ void __launch_bounds__(256)
flux_kernel(const DiffArgs a, const int t, double* __restrict__ flux_east, double* __restrict__ flux_north,
            double* __restrict__ flux_top) {
    extern __shared__ double sm[];
    const int N = a.g.N, M = a.g.M, nz = a.g.nz;
    const int i = blockIdx.y;
    const int p0 = blockIdx.x * blockDim.x;
    const int jlo = p0 / nz;
    const int jhi = min(M - 1, (p0 + (int)blockDim.x - 1) / nz);
    const int jr = jhi - jlo + 1;

    Divisor* d4zt = reinterpret_cast<Divisor*>(sm);  // 4*dzt[k]
    Divisor* dcdxu = d4zt + nz;                      // cost[j]*dxu[i]
    Divisor* ddyu = dcdxu + jr;                      // dyu[j]
    Divisor* dcost = ddyu + jr;                      // cost[j]
    Divisor* d4ytc = dcost + jr;                     // (4*dyt[j])*cost[j]
    Divisor* d4xt = d4ytc + jr;                      // 4*dxt[i] (one entry)
    if (threadIdx.x == 0) d4xt[0] = make_divisor(mul(4.0, a.g.dxt[i]));
    for (int k = threadIdx.x; k < nz; k += blockDim.x) d4zt[k] = make_divisor(mul(4.0, a.g.dzt[k]));
    for (int q = threadIdx.x; q < jr; q += blockDim.x) {
        const int j = jlo + q;
        dcdxu[q] = make_divisor(mul(a.g.cost[j], a.g.dxu[i]));
        ddyu[q] = make_divisor(a.g.dyu[j]);
        dcost[q] = make_divisor(a.g.cost[j]);
        d4ytc[q] = make_divisor(mul(mul(4.0, a.g.dyt[j]), a.g.cost[j]));
    }
    __syncthreads();

    const int p = p0 + threadIdx.x;
    if (p >= M * nz) return;
    const int j = p / nz;
    const int k = p - j * nz;
    const int q = j - jlo;
    const size_t plane = (size_t)M * nz;
    const size_t c = (size_t)i * plane + p;

    const bool inE = (i >= 1 && i < N - 2 && j >= 2 && j < M - 2);
    const bool inN = (i >= 2 && i < N - 2 && j >= 1 && j < M - 2);
    const bool inT = (i >= 2 && i < N - 2 && j >= 2 && j < M - 2 && k < nz - 1);
    double fe = 0.0, fn = 0.0, ft = 0.0;

    if (inE || inN || inT) {
        const int tau = *a.tau;
        const double* __restrict__ tr = a.t[t].tr + tau;
        const double* __restrict__ K = a.K;
        // K1 = K_iso - K_skew, K2 = K_iso + K_skew with the unused one 0.0 (diffusion.py:15-16,180-188)
        auto K1 = [&](size_t cell) { return SKEW ? sub(0.0, __ldg(K + cell)) : __ldg(K + cell); };
        const int km = k > 0 ? -1 : 0, kp = k < nz - 1 ? 1 : 0;  // pad_z_edges clamping
        const double tc = ldt(tr, c), tcm = ldt(tr, c + km), tcp = ldt(tr, c + kp);
        const double dz0_c = sub(tc, tcm), dz1_c = sub(tcp, tc);

        if (inE) {  // diffusion.py:25-47
            const size_t ce = c + plane;
            const double te = ldt(tr, ce), tem = ldt(tr, ce + km), tep = ldt(tr, ce + kp);
            double diffloc;
            if (k > 0)
                diffloc = mul(0.25, add(add(add(K1(c), K1(c - 1)), K1(ce)), K1(ce - 1)));
            else
                diffloc = mul(0.5, add(K1(c), K1(ce)));
            const double2 A0 = __ldg(reinterpret_cast<const double2*>(a.Ai_ez + c * 4));      // ip=0: kr=0,1
            const double2 A1 = __ldg(reinterpret_cast<const double2*>(a.Ai_ez + c * 4) + 1);  // ip=1
            double sumz = add(0.0, mul(mul(diffloc, A0.x), dz0_c));
            sumz = add(sumz, mul(mul(diffloc, A1.x), sub(te, tem)));
            sumz = add(sumz, mul(mul(diffloc, A0.y), dz1_c));
            sumz = add(sumz, mul(mul(diffloc, A1.y), sub(tep, te)));
            fe = add(strict::div(sumz, d4zt[k]), mul(strict::div(sub(te, tc), dcdxu[q]), __ldg(a.K_11 + c)));
        }
        if (inN) {  // diffusion.py:52-77
            const size_t cn = c + nz;
            const double tn = ldt(tr, cn), tnm = ldt(tr, cn + km), tnp = ldt(tr, cn + kp);
            double diffloc;
            if (k > 0)
                diffloc = mul(0.25, add(add(add(K1(c), K1(c - 1)), K1(cn)), K1(cn - 1)));
            else
                diffloc = mul(0.5, add(K1(c), K1(cn)));
            const double2 A0 = __ldg(reinterpret_cast<const double2*>(a.Ai_nz + c * 4));
            const double2 A1 = __ldg(reinterpret_cast<const double2*>(a.Ai_nz + c * 4) + 1);
            double sumz = add(0.0, mul(mul(diffloc, A0.x), dz0_c));
            sumz = add(sumz, mul(mul(diffloc, A1.x), sub(tn, tnm)));
            sumz = add(sumz, mul(mul(diffloc, A0.y), dz1_c));
            sumz = add(sumz, mul(mul(diffloc, A1.y), sub(tnp, tn)));
            fn = mul(__ldg(a.g.cosu + j),
                     add(strict::div(sumz, d4zt[k]), mul(strict::div(sub(tn, tc), ddyu[q]), __ldg(a.K_22 + c))));
        }
        if (inT) {  // diffusion.py:85-111 (k < nz-1 here, so level k+1 exists)
            const size_t ce = c + plane, cw = c - plane, cn = c + nz, cs = c - nz;
            const double diffloc = SKEW ? add(0.0, __ldg(K + c)) : __ldg(K + c);
            const double2 X0 = __ldg(reinterpret_cast<const double2*>(a.Ai_bx + c * 4));
            const double2 X1 = __ldg(reinterpret_cast<const double2*>(a.Ai_bx + c * 4) + 1);
            const double2 Y0 = __ldg(reinterpret_cast<const double2*>(a.Ai_by + c * 4));
            const double2 Y1 = __ldg(reinterpret_cast<const double2*>(a.Ai_by + c * 4) + 1);
            const double tu = tcp;  // (i,j,k+1)
            const double te = ldt(tr, ce), teu = ldt(tr, ce + 1), tw = ldt(tr, cw), twu = ldt(tr, cw + 1);
            const double tn = ldt(tr, cn), tnu = ldt(tr, cn + 1), ts = ldt(tr, cs), tsu = ldt(tr, cs + 1);
            // ip outer, kr inner
            double sumx = add(0.0, mul(strict::div(mul(diffloc, X0.x), dcost[q]), sub(tc, tw)));
            sumx = add(sumx, mul(strict::div(mul(diffloc, X0.y), dcost[q]), sub(tu, twu)));
            sumx = add(sumx, mul(strict::div(mul(diffloc, X1.x), dcost[q]), sub(te, tc)));
            sumx = add(sumx, mul(strict::div(mul(diffloc, X1.y), dcost[q]), sub(teu, tu)));
            const double cu0 = __ldg(a.g.cosu + j - 1), cu1 = __ldg(a.g.cosu + j);
            double sumy = add(0.0, mul(mul(mul(diffloc, Y0.x), cu0), sub(tc, ts)));
            sumy = add(sumy, mul(mul(mul(diffloc, Y0.y), cu0), sub(tu, tsu)));
            sumy = add(sumy, mul(mul(mul(diffloc, Y1.x), cu1), sub(tn, tc)));
            sumy = add(sumy, mul(mul(mul(diffloc, Y1.y), cu1), sub(tnu, tu)));
            ft = add(strict::div(sumx, d4xt[0]), strict::div(sumy, d4ytc[q]));
        }
    }
    flux_east[c] = fe;
    flux_north[c] = fn;
    flux_top[c] = ft;
}

// ------------------------------------------------------------------------------------------------
// update_kernel: grid (tiles over j in [1, M-1), planes i in [1, N-1)).  A tile is `cols` whole
// columns of one x-plane = one contiguous piece of every (N,M,nz) array.  Small CTAs (128 threads,
// ~600 cells) so that several are resident per SM and the latency-bound column solves of one
// overlap the streaming phases of the others.
// Shared memory (doubles, odd column pitch): L, D, U, R[NTR] + metric tables.
//   phase B  per cell: explicit flux divergence, tracer + tendency update, right-hand sides, matrix
//   phase C  per column: dgtsv elimination on NTR right-hand sides (tdma_device.cuh)
//   phase D  per cell: implicit result, tendency, dissipation (the T-point dissipation of levels
//            k and k+1 is recomputed from the fluxes instead of being staged)
// ------------------------------------------------------------------------------------------------
}  // namespace
}  // namespace vb

#include "iso_update_phases.cuh"

namespace vb {
using namespace upd;
namespace {

// ---- one tile per CTA, phases separated by CTA barriers --------------------------------------------------
// (A persistent, warp-specialised variant -- warp 0 solving tile n while warps 1-3 stream tiles n+1 and
// n-1 through a second shared-memory buffer, named barriers FULL/DONE -- was built on these same phase
// functions, verified bit-identical and measured 8-10 % SLOWER on both benchmark grids: with the tile
// double buffered only half as many columns are being solved per SM, and 12 streaming warps per SM move
// less data than the 24 of this kernel.  DESIGN.md section 3.2.)
template <int NTR, bool SKEW, bool ENERGY, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 768 / BLOCK)
update_kernel(const DiffArgs a, const Scratch f, const int cols, const int pitch, const double fac_diss,
              const double gr) {
    extern __shared__ double sm[];
    const int N = a.g.N, nz = a.g.nz;
    const int i = 1 + blockIdx.y;
    if ((i == 1 && a.skip_west_ring) || (i == N - 2 && a.skip_east_ring)) return;  // a neighbouring sub-slab's interior
    const UpdConst u = upd_const(a, pitch, fac_diss, gr);
    LevelTabs lv;
    lv.ddzt = reinterpret_cast<Divisor*>(sm);
    lv.ddzw = lv.ddzt + nz;
    lv.dt_dzw = reinterpret_cast<double*>(lv.ddzw + nz);
    const TileBuf b = tile_buf_at(lv.dt_dzw + nz, cols, pitch, NTR);
    const TileGeom g = tile_geom(u, i, 1 + blockIdx.x * cols, cols);
    fill_level_tabs(a, lv, nz, u.dt, threadIdx.x, BLOCK);
    fill_tile_tabs(a, u, b, g, threadIdx.x, BLOCK);
    __syncthreads();
    phase_b<NTR, SKEW, ENERGY, false>(a, f, u, lv, b, g, threadIdx.x, BLOCK);
    __syncthreads();  // also orders this CTA's diss[] stores before the phase D loads of its own cells
    if (!SKEW) {
        phase_c<NTR>(u, b, g, threadIdx.x, BLOCK);
        __syncthreads();
    }
    phase_d<NTR, SKEW, ENERGY, false>(a, f, u, lv, b, g, threadIdx.x, BLOCK);
}

template <int NTR, bool SKEW, bool ENERGY>
void launch_update(cudaStream_t s, const DiffArgs& a, const Scratch& f) {
    const int M = a.g.M, N = a.g.N, nz = a.g.nz;
    const int pitch = (nz + 1) | 1;  // odd (no bank conflicts) and > nz: the element past a column belongs to nobody
    const double fac_diss = 0.5 * a.grav / a.rho_0;  // diffusion.py (core) :19-21, Python float arithmetic
    const double gr = -a.grav / a.rho_0;             // isoneutral/diffusion.py:259,268
    int cols = max(1, 640 / nz);
    if (const char* e = getenv("VEROS_B200_UPD_CELLS")) cols = max(1, atoi(e) / nz);  // tuning knob (tile cells)
    cols = min(cols, M - 2);
    const size_t lev_bytes = (size_t)nz * (2 * sizeof(Divisor) + 8);
    const int tiles_per_row = (M - 2 + cols - 1) / cols;
    const int want_tiles = 4 * 148;  // small grids: spread over the SMs
    const int rows = N - 2;
    if (tiles_per_row * rows < want_tiles) {
        const int per_row = (want_tiles + rows - 1) / rows;
        cols = max(1, (M - 2 + per_row - 1) / per_row);
    }
    const size_t smem = lev_bytes + 8 * ((size_t)(3 + NTR) * cols * pitch + (size_t)cols * 4 + (cols + 1) / 2 + 1) + 16;
    dim3 grid((M - 2 + cols - 1) / cols, N - 2);
    static const int block = getenv("VEROS_B200_UPD_BLOCK") ? atoi(getenv("VEROS_B200_UPD_BLOCK")) : 128;  // tuning knob
    if (block == 256) {
        auto kern = update_kernel<NTR, SKEW, ENERGY, 256>;
        allow_big_smem(kern, 200 * 1024);
        kern<<<grid, 256, smem, s>>>(a, f, cols, pitch, fac_diss, gr);
    } else {
        auto kern = update_kernel<NTR, SKEW, ENERGY, 128>;
        allow_big_smem(kern, 200 * 1024);
        kern<<<grid, 128, smem, s>>>(a, f, cols, pitch, fac_diss, gr);
    }
    count_launch();
    check_launch("update_kernel");
}

}  // namespace

// workspace layout (doubles): per tracer flux_east, flux_north, flux_top (3*ntr arrays of N*M*nz),
// then one dissipation array per tracer
size_t diffusion_workspace_doubles(int N, int M, int nz, int ntr) { return (size_t)4 * ntr * N * M * nz; }

void launch_iso_diffusion_ws(cudaStream_t s, const DiffArgs& a, double* ws) {
    const int N = a.g.N, M = a.g.M, nz = a.g.nz;
    if (N < 5 || M < 5 || nz < 2) return;
    const size_t n3 = (size_t)N * M * nz;
    Scratch f;
    f.diss[0] = ws + (size_t)(3 * a.ntr) * n3;
    f.diss[1] = f.diss[0] + (a.ntr > 1 ? n3 : 0);
    for (int t = 0; t < a.ntr; ++t) {
        double* fe = ws + (size_t)(3 * t) * n3;
        double* fn = fe + n3;
        double* ft = fn + n3;
        f.fe[t] = fe;
        f.fn[t] = fn;
        f.ft[t] = ft;
        if (a.fluxes_ready) continue;
        const int block = 256;
        dim3 grid((M * nz + block - 1) / block, N);
        const int jr = block / nz + 2;
        const size_t smem = sizeof(Divisor) * ((size_t)nz + 4 * jr + 1);
        if (a.skew)
            flux_kernel<true><<<grid, block, smem, s>>>(a, t, fe, fn, ft);
        else
            flux_kernel<false><<<grid, block, smem, s>>>(a, t, fe, fn, ft);
        count_launch();
        if (!check_launch("flux_kernel")) return;
    }
    if (a.ntr == 1) {
        f.fe[1] = f.fe[0]; f.fn[1] = f.fn[0]; f.ft[1] = f.ft[0];
    }
#define VB_DISPATCH(NTR)                                                                   \
    if (a.skew) {                                                                          \
        if (a.energy) launch_update<NTR, true, true>(s, a, f);                             \
        else launch_update<NTR, true, false>(s, a, f);                                     \
    } else {                                                                               \
        if (a.energy) launch_update<NTR, false, true>(s, a, f);                            \
        else launch_update<NTR, false, false>(s, a, f);                                    \
    }
    if (a.ntr == 2) { VB_DISPATCH(2) } else { VB_DISPATCH(1) }
#undef VB_DISPATCH
}

}  // namespace vb
