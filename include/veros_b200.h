/*
 * veros_b200.h -- C ABI of libveros_b200.so: sm_100a CUDA implementation of Veros's isoneutral
 * mixing step and column (tridiagonal) solve.
 *
 * Every compute entry point has the signature of a *legacy XLA GPU custom call* (api_version 0),
 * the mechanism the reference uses for its one native op:
 *
 *     void f(cudaStream_t stream, void** buffers, const char* opaque, size_t opaque_len);
 *
 *   reference declaration   veros/core/special/cuda_tdma_kernels.h:10-11
 *   reference registration  veros/core/special/tdma_.py:34-37
 *                           jax.ffi.register_ffi_target(name, PyCapsule("xla._CUSTOM_CALL_TARGET"),
 *                                                       platform="CUDA", api_version=0)
 *   reference capsule       veros/core/special/tdma_cuda_.pyx:23-30
 *
 * `buffers` holds device pointers: all operands in order, then all results in order.  In/out state
 * (Ai_*, K_*, tracers, tendencies, P_diss) appears once as an operand and once as a result; the
 * intended use is operand_output_aliases (same pointer twice).  If the two pointers differ the op
 * first copies operand -> result on `stream`, so it is also correct without aliasing.
 * `opaque` is the raw bytes of one of the POD descriptors below (length-checked).
 * A trailing 5th argument (XlaCustomCallStatus*) passed by newer XLA versions is ignored.
 *
 * The ops only ENQUEUE work on `stream`: no allocation, no synchronisation, no retained state.
 * Errors (bad descriptor, failed launch) never throw or exit: the first one since the latch was last cleared is
 * printed to stderr and latched; read it with veros_b200_last_error().  A call stops enqueueing after its own
 * first failure; an error latched by an earlier call does not turn later calls into no-ops.  (The reference exits the process instead,
 * cuda_tdma_kernels.cu:9-17,72-79.)
 *
 * Array conventions (veros/variables.py:75-162): C order, z fastest; N = nx+4, M = ny+4 include the
 * two ghost cells per side; float64; masks are 1-byte bool; kbot/tau/taup1 are int32.
 * Tracers are (N,M,nz,3) with the time level last; Ai_* are (N,M,nz,2,2).
 */
#ifndef VEROS_B200_H
#define VEROS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VEROS_B200_ABI_VERSION 1

/* ---------------------------------------------------------------- descriptors (opaque payloads) */

/* Same layout as the reference's TridiagDescriptor (cuda_tdma_kernels.h:5-8,
 * tdma_cuda_.pyx:18-20 build_tridiag_descriptor). */
typedef struct VerosB200TridiagDescriptor {
    int32_t num_systems;  /* number of columns = prod(shape[:-1]) */
    int32_t system_depth; /* nz */
} VerosB200TridiagDescriptor;

#define VEROS_B200_HAS_B_EDGE 1
#define VEROS_B200_HAS_D_EDGE 2

typedef struct VerosB200SolveDescriptor {
    int32_t num_systems;
    int32_t system_depth;
    int32_t flags; /* VEROS_B200_HAS_B_EDGE | VEROS_B200_HAS_D_EDGE */
    int32_t reserved;
} VerosB200SolveDescriptor;

#define VEROS_B200_FLAG_SKEW 1 /* iso_diffusion: isoneutral_skew_diffusion (K1=-K_gm, K2=K_gm) */
/* The arrays passed are an x-sub-slab (planes [i0, i1) of a wider local slab, via pointer offsets -- x is
 * the slowest axis, so a sub-slab is contiguous).  Every output is idempotent under overlapping sub-slabs
 * except the P_diss accumulation on the one-cell ring [1:-1] outside the interior (veros/core/diffusion.py:
 * 15-35,41-62): with these flags the west (plane 1) / east (plane N-2) ring plane is left to the neighbouring
 * sub-slab, whose interior it is.  Used to pipeline host<->device copies and to overlap the halo exchange. */
#define VEROS_B200_FLAG_NO_WEST_RING 2
#define VEROS_B200_FLAG_NO_EAST_RING 4
/* Tuning / test knobs of the slope kernel: force the one-launch (all faces) or the two-launch (east+north, then
 * top faces) instantiation instead of choosing by grid size.  Results are identical either way. */
#define VEROS_B200_FLAG_PRE_SINGLE 8
#define VEROS_B200_FLAG_PRE_SPLIT 16
/* Masked cells.  By default the kernels do not compute what the masks force to zero: a face whose maskU / maskV /
 * maskW is 0 only stores the zeros the reference's expressions produce there (exact for finite fields, any masks),
 * and the fused step leaves dry cells (maskT = 0) of the tracers, tendencies and P_diss_iso alone, which equals the
 * reference bit for bit when the masks are the ones Veros derives from kbot (veros/core/numerics.py:200-221: maskT from
 * kbot, maskU/V/W the staggered minima).  Warps that lie entirely below the sea floor or on land then cost nothing.
 * This flag turns the shortcut off (every cell is computed; for inputs with inconsistent masks or non-finite values
 * on land). */
#define VEROS_B200_FLAG_NO_MASK_SKIP 64
/* iso_step: run the whole step as ONE persistent kernel (csrc/iso_mega.cu: all phases as dependency-ordered work
 * items, inter-phase scratch in an L2-resident ring) instead of separate launches (table setup, TEOS-10 pass, slope +
 * flux kernel(s), update kernel).  Same results bit for bit.  Opt-in: measured slower than the separate launches on
 * every grid, kept with its parity tests and measurements. */
#define VEROS_B200_FLAG_STEP_FUSED 32

/* Static (jit-constant) facts of the isoneutral ops: shapes and the settings of
 * veros/settings.py:24-91 that the path reads. */
typedef struct VerosB200IsoDescriptor {
    int32_t nx_tot; /* N = nx + 4 (local slab, ghosts included) */
    int32_t ny_tot; /* M = ny + 4 */
    int32_t nz;
    int32_t eq_of_state_type;       /* 1..5, veros/core/density/get_rho.py:93-131 */
    int32_t enable_conserve_energy; /* settings.enable_conserve_energy */
    int32_t flags;                  /* VEROS_B200_FLAG_* */
    double K_iso_steep;
    double iso_slopec;
    double iso_dslope;
    double dt_tracer;
    double grav;
    double rho_0;
} VerosB200IsoDescriptor;

/* vertmix_tempsalt: sizes and the tracer time step. */
typedef struct VerosB200VmixDescriptor {
    int32_t nx_tot; /* N = nx + 4 */
    int32_t ny_tot; /* M = ny + 4 */
    int32_t nz;
    int32_t flags; /* reserved, 0 */
    double dt_tracer;
} VerosB200VmixDescriptor;

/* implicit_vert_friction and isoneutral_diag_streamfunction: sizes (and the momentum time step). */
typedef struct VerosB200ColumnDescriptor {
    int32_t nx_tot; /* N = nx + 4 */
    int32_t ny_tot; /* M = ny + 4 */
    int32_t nz;
    int32_t flags; /* reserved, 0 */
    double dt;     /* dt_mom for implicit_vert_friction; unused otherwise */
} VerosB200ColumnDescriptor;

/* set_eke_diffusivities: sizes and the settings of veros/settings.py that veros/core/eke.py:34-85 reads. */
typedef struct VerosB200EkeDescriptor {
    int32_t nx_tot, ny_tot, nz;
    int32_t enable_eke;
    int32_t enable_eke_isopycnal_diffusion;
    int32_t flags; /* reserved, 0 */
    double pi, eke_lmin, eke_cross, eke_crhin, eke_k_max, eke_c_k, K_gm_0, K_iso_0;
} VerosB200EkeDescriptor;

/* advect_tempsalt: sizes and the settings veros/core/thermodynamics.py:10-62,223-245 reads. */
#define VEROS_B200_ADVECT_SUPERBEE 1 /* settings.enable_superbee_advection */
#define VEROS_B200_ADVECT_NO_AB 2    /* tendencies only (advect_temperature / advect_salinity), no Adams-Bashforth step */
typedef struct VerosB200AdvectDescriptor {
    int32_t nx_tot, ny_tot, nz;
    int32_t flags; /* VEROS_B200_ADVECT_* */
    double dt_tracer;
    double AB_eps;
} VerosB200AdvectDescriptor;

/* ------------------------------------------------------------------------------ compute ops */

/* Column solve on the model's native (X,Y,nz) z-contiguous layout, replacing
 * veros.core.utilities.solve_implicit (utilities.py:51-59) + operators.solve_tridiagonal
 * (operators.py:60-77 NumPy/dgtsv, :103-156 JAX) + special.tdma_ (tdma_.py:49-68).
 * Per column it executes LAPACK dgtsv's operation sequence (partial pivoting included), i.e. it is
 * bit-identical to the reference's NumPy backend; rows outside water_mask give 0.
 * opaque: VerosB200SolveDescriptor.
 * buffers: 0 a, 1 b, 2 c, 3 d (f64 [ncol][nz]), 4 water_mask, 5 edge_mask (u8 [ncol][nz]),
 *          6 b_edge, 7 d_edge (f64; ignored unless flagged, pass any valid pointer) | 8 out (f64). */
void veros_b200_solve_implicit_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* Drop-in for the reference's own CUDA targets "tdma_cuda_double" / "tdma_cuda_float"
 * (cuda_tdma_kernels.cu:81-111): pre-masked diagonals in the z-major layout requested by
 * tdma_.py:167-180, Thomas recurrence without pivoting.
 * opaque: VerosB200TridiagDescriptor.  buffers: 0 a, 1 b, 2 c, 3 d | 4 out, 5 workspace. */
void veros_b200_tdma_zmajor_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);
void veros_b200_tdma_zmajor_f32(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* veros.core.isoneutral.isoneutral_diffusion_pre (isoneutral/isoneutral.py:18-229).
 * opaque: VerosB200IsoDescriptor.
 * buffers (operands): 0 temp, 1 salt (N,M,nz,3), 2 tau (int32[1]), 3 K_iso,
 *          4 maskT, 5 maskU, 6 maskV, 7 maskW (u8), 8 dxt, 9 dxu (N), 10 dyt, 11 dyu, 12 cost,
 *          13 cosu (M), 14 dzt, 15 dzw, 16 zt (nz),
 *          17 Ai_ez, 18 Ai_nz, 19 Ai_bx, 20 Ai_by, 21 K_11, 22 K_22, 23 K_33   (previous values)
 * (results): 24 Ai_ez, 25 Ai_nz, 26 Ai_bx, 27 Ai_by, 28 K_11, 29 K_22, 30 K_33,
 *          31 workspace (veros_b200_iso_pre_workspace_bytes) */
void veros_b200_iso_pre_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* veros.core.isoneutral.isoneutral_diffusion / isoneutral_skew_diffusion for ONE tracer
 * (isoneutral/diffusion.py:9-307; veros/core/diffusion.py:9-62 for the dissipation).
 * opaque: VerosB200IsoDescriptor (flags: VEROS_B200_FLAG_SKEW selects the skew variant).
 * buffers (operands): 0 tr (N,M,nz,3), 1 dtracer_iso, 2 P_diss (iso: P_diss_iso, skew: P_diss_skew),
 *          3 tau, 4 taup1 (int32[1]), 5 K (iso: K_iso, skew: K_gm),
 *          6 Ai_ez, 7 Ai_nz, 8 Ai_bx, 9 Ai_by, 10 K_11, 11 K_22, 12 K_33, 13 maskT, 14 maskW (u8),
 *          15 kbot (int32 N,M), 16 dxt, 17 dxu, 18 dyt, 19 dyu, 20 cost, 21 cosu, 22 dzt, 23 dzw,
 *          24 int_drhodX (N,M,nz,3; int_drhodT for temp, int_drhodS for salt)
 * (results): 25 tr, 26 dtracer_iso, 27 P_diss, 28 workspace (veros_b200_iso_diffusion_workspace_bytes)
 * With enable_conserve_energy == 0 buffers 2, 24, 27 are not touched (pass any valid pointer).
 * Bit-identical to the reference's NumPy backend on identical inputs. */
void veros_b200_iso_diffusion_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* One complete neutral-diffusion step, veros/core/thermodynamics.py:430-432:
 *   isoneutral_diffusion_pre; isoneutral_diffusion(temp); isoneutral_diffusion(salt)
 * with the two tracer updates sharing one factorisation of the K_33 column matrix.
 * opaque: VerosB200IsoDescriptor.
 * buffers (operands): 0 temp, 1 salt, 2 dtemp_iso, 3 dsalt_iso, 4 P_diss_iso,
 *          5 Ai_ez, 6 Ai_nz, 7 Ai_bx, 8 Ai_by, 9 K_11, 10 K_22, 11 K_33,
 *          12 tau, 13 taup1, 14 K_iso, 15 maskT, 16 maskU, 17 maskV, 18 maskW, 19 kbot,
 *          20 dxt, 21 dxu, 22 dyt, 23 dyu, 24 cost, 25 cosu, 26 dzt, 27 dzw, 28 zt,
 *          29 int_drhodT, 30 int_drhodS
 * (results): 31..42 = operands 0..11, 43 workspace (veros_b200_iso_step_workspace_bytes) */
void veros_b200_iso_step_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* The step right after the isoneutral path: vertmix_tempsalt,
 * veros/core/thermodynamics.py:248-288 -- coefficient assembly from kappaH, the two solve_implicit calls
 * (one dgtsv factorisation shared by both tracers) and the dtemp_vmix / dsalt_vmix tendencies in one
 * kernel; bit-identical to the reference's NumPy backend.  The enforce_boundaries calls that follow in
 * the reference (:290-297) are the caller's halo exchange (veros_b200_halo_pack_unpack + NCCL).
 * opaque: VerosB200VmixDescriptor.
 * buffers (operands): 0 temp, 1 salt (N,M,nz,3), 2 taup1 (int32[1]), 3 kappaH (N,M,nz),
 *          4 forc_temp_surface, 5 forc_salt_surface (N,M), 6 kbot (int32 N,M), 7 dzt, 8 dzw (nz)
 * (results): 9 temp, 10 salt, 11 dtemp_vmix, 12 dsalt_vmix (N,M,nz; every element written) */
void veros_b200_vertmix_tempsalt_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* SURVEY.md 8f rank 3 -- the first other caller of solve_implicit with its coefficient assembly fused in:
 * implicit_vert_friction, veros/core/friction.py:92-205 (both velocity components, du_mix / dv_mix, the dissipation
 * through ugrid_to_tgrid / vgrid_to_tgrid, numerics.py:313-336, added to K_diss_v).  Bit-identical to the NumPy backend.
 * opaque: VerosB200ColumnDescriptor (dt = dt_mom).
 * buffers (operands): 0 u, 1 v (N,M,nz,3), 2 du_mix, 3 dv_mix, 4 K_diss_v (N,M,nz), 5 tau, 6 taup1 (int32[1]),
 *          7 kappaM (N,M,nz), 8 maskU, 9 maskV (u8), 10 kbot (int32 N,M), 11 dzt, 12 dzw (nz), 13 dxt, 14 dxu (N),
 *          15 area_v, 16 area_t (N,M)
 * (results): 17..21 = operands 0..4, 22 workspace (2 * N*M*nz doubles) */
void veros_b200_implicit_vert_friction_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* SURVEY.md 8f rank 4, consumer of the path's outputs: isoneutral_diag_streamfunction_kernel,
 * veros/core/isoneutral/isoneutral.py:232-258.  opaque: VerosB200ColumnDescriptor.
 * buffers (operands): 0 K_gm, 1 Ai_ez, 2 Ai_nz, 3 B1_gm, 4 B2_gm | (results): 5 B1_gm, 6 B2_gm */
void veros_b200_iso_diag_streamfunction_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* SURVEY.md 8f rank 4, producer of K_gm / K_iso: set_eke_diffusivities_kernel, veros/core/eke.py:34-85 (NumPy's
 * pairwise order of the Rossby-radius column sum included).  opaque: VerosB200EkeDescriptor.
 * buffers (operands): 0 Nsqr, 1 eke (N,M,nz,3), 2 tau (int32[1]), 3 maskW (u8), 4 dzw (nz), 5 coriolis_t, 6 beta (N,M)
 * (results): 7 L_rossby (N,M), 8 L_rhines, 9 eke_len, 10 sqrteke, 11 K_gm, 12 K_iso (N,M,nz)
 * With enable_eke == 0 only K_gm and K_iso are written and buffers 0..10 are not touched. */
void veros_b200_set_eke_diffusivities_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* SURVEY.md 8f rank 4, producer of temp/salt[..., taup1]: advect_temperature + advect_salinity
 * (veros/core/thermodynamics.py:43-62 -> advect_tracer :10-40 with adv_flux_2nd / adv_flux_superbee,
 * veros/core/advection.py:8-115) and the Adams-Bashforth step (:223-245), one pass for both tracers.
 * opaque: VerosB200AdvectDescriptor.
 * buffers (operands): 0 temp, 1 salt, 2 dtemp, 3 dsalt (N,M,nz,3), 4 tau, 5 taup1, 6 taum1 (int32[1]),
 *          7 u, 8 v, 9 w (N,M,nz,3), 10 maskT, 11 maskU, 12 maskV, 13 maskW (u8), 14 dxt (N), 15 dyt (M), 16 dzt (nz),
 *          17 cost, 18 cosu (M)
 * (results): 19 temp, 20 salt, 21 dtemp, 22 dsalt = operands 0..3 */
void veros_b200_advect_tempsalt_f64(void* stream, void** buffers, const char* opaque, size_t opaque_len);

/* ------------------------------------------------------------------------ host-side helpers */

/* Scratch the caller must provide as the last result (0 is possible; then pass any valid pointer). */
size_t veros_b200_iso_pre_workspace_bytes(const char* opaque, size_t opaque_len);
size_t veros_b200_iso_diffusion_workspace_bytes(const char* opaque, size_t opaque_len);
size_t veros_b200_iso_step_workspace_bytes(const char* opaque, size_t opaque_len);

/* Byte offset, inside the iso_step workspace, of 16 uint64 scheduling statistics the fused step kernel leaves behind
 * (development aid): [t] ns its CTAs spent waiting for dependencies before items of type t, [5+t] ns spent working on
 * them, [10+t] number of items; t = 1 TEOS-10, 2 staging copy, 3 slopes+fluxes, 4 update. */
size_t veros_b200_iso_step_stats_offset(const char* opaque, size_t opaque_len);

/* Sticky error state: 0 = ok; otherwise a cudaError_t value or VEROS_B200_ERR_*. */
#define VEROS_B200_ERR_BAD_DESCRIPTOR 100001
#define VEROS_B200_ERR_BAD_ARGUMENT 100002
int veros_b200_last_error(void);
const char* veros_b200_last_error_string(void);
void veros_b200_clear_error(void);

/* Multi-GPU plumbing (not a custom call): gather (mode 0) the west/east interior edge planes [2,4) and
 * [N-4,N-2) of up to 4 fields into contiguous buffers, or scatter (mode 1) received buffers into the ghost
 * planes [0,2) and [N-2,N).  Fields are (N,M,nz[,nlev]) float64; `level` selects the time level of tracers
 * (nlev = 3) -- what veros/core/thermodynamics.py:293-298 exchanges.  A NULL buffer skips that side.
 * Buffer layout: field-major, 2*M*nz doubles per field. */
void veros_b200_halo_pack_unpack(void* stream, int mode, void** fields, int nfields, int nx_tot, int ny_tot, int nz,
                                 int nlev, int level, void* west_buf, void* east_buf);

/* Multi-GPU plumbing over peer memory (not a custom call): the same exchange without staging buffers or NCCL.
 * `west_fields` / `east_fields` are the neighbours' arrays and `west_flags` / `east_flags` their flag words, mapped
 * into this process through CUDA IPC (NULL flags: no neighbour on that side; a single rank on a ring passes its
 * own pointers).  The kernel announces that this rank's ghost planes may be overwritten, waits for the neighbours
 * to say the same, stores this rank's edge planes [2,4) / [N-4,N-2) into the neighbours' ghost planes, publishes
 * completion and returns when both neighbours have published theirs.  `seq` = 1, 2, 3, ... counts the exchanges of
 * this set of fields (same value on all ranks); `my_flags` = 4 int32 zeroed once, `counter` = 1 uint32 zeroed once.
 * Every participating rank must enqueue the call; a rank enqueues nothing that needs the exchange before it. */
void veros_b200_halo_put(void* stream, int seq, void** fields, void** west_fields, void** east_fields, int nfields,
                         int nx_tot, int nx_tot_west, int ny_tot, int nz, int nlev, int level, void* my_flags,
                         void* west_flags, void* east_flags, void* counter);

/* CUDA IPC helpers for veros_b200_halo_put.  get_handle: the 64-byte cudaIpcMemHandle_t of the cudaMalloc block
 * starting at `base` (0 on success).  open_handle: maps a neighbour's block with `device` -- the importing rank's
 * compute device -- current, peer access enabled lazily; returns the mapped base or NULL (error latched).
 * close: unmaps. */
int veros_b200_ipc_get_handle(void* base, void* handle64);
void* veros_b200_ipc_open_handle(int device, const void* handle64);
void veros_b200_ipc_close(void* base);

/* Measurement hook: while set (n >= 4, events created by the caller), veros_b200_iso_step_f64 records
 * events[0] on its stream before its first kernel, [1] before the slope+flux kernel, [2] after it and
 * [3] after the update kernel, so a benchmark can time the dominant kernel inside the fused call with
 * CUDA events.  Pass NULL / 0 to clear.  Not for concurrent use from several threads. */
void veros_b200_profile_events(void** events, int n);

int veros_b200_abi_version(void);
/* sizeof() of descriptor 0 = Tridiag, 1 = Solve, 2 = Iso, 3 = Vmix, 4 = Column, 5 = Eke, 6 = Advect as compiled into the library. */
size_t veros_b200_descriptor_size(int which);
/* Number of kernel launches enqueued by this library since load (bench.py's gpu_launches). */
unsigned long long veros_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VEROS_B200_H */
