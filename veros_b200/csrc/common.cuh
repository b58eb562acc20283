// Shared declarations of libveros_b200: kernel launch wrappers, error latch, index helpers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/veros_b200.h"

namespace vb {

// ---- error latch (api.cu) -------------------------------------------------------------------
void set_error(int code, const char* what);
void count_launch(int n = 1);
bool check_launch(const char* what);  // cudaPeekAtLastError -> latch; true if ok
bool call_failed();                   // an error was latched by the current entry-point call on this thread
void begin_call();
constexpr int kMaxDevices = 64;        // per-device caches (occupancy, function attributes) are indexed by device
// Opt a kernel in to more than 48 KB of dynamic shared memory on the CURRENT device.  The attribute is per
// (kernel, device), so the "already done" record is keyed by both (a process may drive several GPUs); after the
// first call for a pair this is a short table lookup, which also keeps CUDA-graph capture free of attribute calls.
inline void allow_big_smem_fn(const void* fn, int bytes) {
    struct Done { const void* fn; int dev, bytes; };
    static std::mutex mu;
    static std::vector<Done> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (const Done& d : done)
        if (d.fn == fn && d.dev == dev && d.bytes >= bytes) return;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    done.push_back(Done{fn, dev, bytes});
}
template <typename K>
inline void allow_big_smem(K kernel, int bytes) { allow_big_smem_fn(reinterpret_cast<const void*>(kernel), bytes); }

// ---- pointers of one isoneutral problem (all device memory) -------------------------------------
struct Grid {
    int N, M, nz;
    const double *dxt, *dxu;          // (N)
    const double *dyt, *dyu, *cost, *cosu;  // (M)
    const double *dzt, *dzw, *zt;     // (nz)
};

struct PreArgs {
    Grid g;
    const double *temp, *salt;  // (N,M,nz,3)
    const int32_t* tau;
    const double* K_iso;
    const uint8_t *maskT, *maskU, *maskV, *maskW;
    double *Ai_ez, *Ai_nz, *Ai_bx, *Ai_by, *K_11, *K_22, *K_33;
    double *drdT, *drdS;  // workspace, EOS 5 only
    double* tables;       // workspace, pre_tables_doubles() doubles: metric tables built per call
    int tables_ready;     // the caller already ran launch_setup_tables on `tables`
    double dt_tracer;
    double two_rd, m2c0, s_max;  // taper constants (filled by launch_iso_pre)
    double* flux[2][3];   // [temp|salt][east|north|top] outputs when with_flux
    int with_flux;
    // fused step with energy: contiguous copies of int_drhodT/S[..., tau] for the update kernel, which is bound
    // by the bandwidth its 24-byte-stride reads waste; made by the slope kernel's CTAs between compute chunks
    const double* stage_src[2];
    double* stage[2];
    int with_stage;
    int no_mask_skip;  // compute masked faces too instead of storing their zeros (VEROS_B200_FLAG_NO_MASK_SKIP)
    int variant;  // 0: by size, 1: one launch for all faces, 2: east+north / top launches (VEROS_B200_FLAG_PRE_*)
    int eos;
    double K_iso_steep, iso_slopec, iso_dslope;
};

struct DiffTracer {
    double* tr;              // (N,M,nz,3) in/out (taup1 level written)
    double* dtracer;         // (N,M,nz) in/out
    const double* int_drhodX;  // (N,M,nz,3)
};

struct DiffArgs {
    Grid g;
    DiffTracer t[2];
    int ntr;  // 1 or 2 tracers sharing the K_33 matrix
    double* P_diss;
    const int32_t *tau, *taup1;
    const double* K;  // K_iso (iso) or K_gm (skew)
    const double *Ai_ez, *Ai_nz, *Ai_bx, *Ai_by, *K_11, *K_22, *K_33;
    const uint8_t *maskT, *maskW;
    const int32_t* kbot;
    int skew, energy;
    int skip_west_ring, skip_east_ring;  // sub-slab mode (VEROS_B200_FLAG_NO_*_RING)
    int fluxes_ready;  // the fused slope+flux kernel already filled the flux workspace
    int dry_skip;      // fused step: cells with maskT = 0 are left alone (see iso_update_phases.cuh)
    const double* stage_x[2];   // contiguous copies of int_drhodX[..., tau] made by the slope kernel (or null)
    double* tables;    // metric tables (tables.cuh), built by launch_setup_tables
    double dt_tracer, grav, rho_0;
};

struct VmixArgs {  // vertmix_tempsalt, veros/core/thermodynamics.py:248-288
    int N, M, nz;
    double *temp, *salt;              // (N,M,nz,3) in/out (taup1 level)
    const int32_t* taup1;
    const double* kappaH;             // (N,M,nz)
    const double *forc_temp, *forc_salt;  // (N,M)
    const int32_t* kbot;
    const double *dzt, *dzw;
    double *dtemp_vmix, *dsalt_vmix;  // (N,M,nz) out
    double dt_tracer;
};

// ---- launchers (one per translation unit) -----------------------------------------------------
void launch_iso_pre(cudaStream_t s, const PreArgs& a, bool profile = false);
void prof_mark(cudaStream_t s, int q);  // api.cu: records the q-th profiling event if a benchmark installed some
size_t diffusion_workspace_doubles(int N, int M, int nz, int ntr);
void launch_iso_diffusion_ws(cudaStream_t s, const DiffArgs& a, double* workspace);
void launch_vertmix(cudaStream_t s, const VmixArgs& a);
// neighbours of the path (next_ops.cu); B = the custom call's buffer list
void launch_implicit_vert_friction(cudaStream_t s, int N, int M, int nz, double dt_mom, void** B);
void launch_diag_streamfunction(cudaStream_t s, int N, int M, int nz, void** B);
void launch_set_eke_diffusivities(cudaStream_t s, const VerosB200EkeDescriptor* d, void** B);
void launch_advect_tempsalt(cudaStream_t s, const VerosB200AdvectDescriptor* d, void** B);  // advect.cu
// the fused persistent step kernel (iso_mega.cu)
size_t mega_ring_doubles(int N, int M, int nz, int eos, int energy);
size_t mega_sync_doubles(int N);
size_t mega_stats_offset_doubles(int N);
void launch_iso_mega(cudaStream_t s, const PreArgs& p, const DiffArgs& d, double* ring, unsigned int* sync);
void launch_solve_implicit(cudaStream_t s, int ncol, int nz, const double* a, const double* b, const double* c,
                           const double* d, const uint8_t* water, const uint8_t* edge, const double* b_edge,
                           const double* d_edge, double* out);
void launch_tdma_zmajor_f64(cudaStream_t s, int nsys, int depth, const double* a, const double* b, const double* c,
                            const double* d, double* out, double* work);
void launch_tdma_zmajor_f32(cudaStream_t s, int nsys, int depth, const float* a, const float* b, const float* c,
                            const float* d, float* out, float* work);

}  // namespace vb
