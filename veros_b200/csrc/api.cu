// extern "C" entry points of libveros_b200.so (see include/veros_b200.h).
// Each one unpacks the XLA custom-call buffer list, validates the opaque descriptor and enqueues
// kernels on the caller's stream.  Nothing here allocates, synchronises or throws.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "tables.cuh"

namespace vb {

static std::atomic<int> g_err{0};
static std::atomic<unsigned long long> g_launches{0};
static std::mutex g_err_mutex;
static char g_err_text[256] = "";
// Set by set_error on the calling thread: an entry point stops enqueueing after ITS OWN first failure, never because
// of an error some other call or thread latched earlier (the latch itself is sticky until the host reads it).
static thread_local bool t_call_failed = false;

void set_error(int code, const char* what) {
    t_call_failed = true;
    std::lock_guard<std::mutex> lock(g_err_mutex);
    if (g_err.load() == 0) {  // first error since the host last cleared the latch: keep its text, print it once
        g_err.store(code);
        std::snprintf(g_err_text, sizeof(g_err_text), "%s (code %d%s%s)", what, code,
                      code < 100000 ? ": " : "", code < 100000 ? cudaGetErrorString((cudaError_t)code) : "");
        std::fprintf(stderr, "veros_b200: %s\n", g_err_text);
    }
}

bool call_failed() { return t_call_failed; }
void begin_call() { t_call_failed = false; }

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }

bool check_launch(const char* what) {
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error((int)e, what);
        return false;
    }
    return true;
}

template <typename T>
static const T* unpack(const char* opaque, size_t len, const char* who) {
    if (opaque == nullptr || len != sizeof(T)) {
        set_error(VEROS_B200_ERR_BAD_DESCRIPTOR, who);
        return nullptr;
    }
    return reinterpret_cast<const T*>(opaque);
}

static bool valid_iso(const VerosB200IsoDescriptor* d, const char* who) {
    if (!d) return false;
    if (d->nx_tot < 5 || d->ny_tot < 5 || d->nz < 2 || d->eq_of_state_type < 1 || d->eq_of_state_type > 5 ||
        (size_t)d->nx_tot * d->ny_tot * d->nz > (size_t)1 << 31 || !(d->iso_dslope > 0.0) || !(d->dt_tracer != 0.0) ||
        !(d->iso_slopec >= 0.0) || !(d->iso_slopec / d->iso_dslope <= 300.0)) {
        set_error(VEROS_B200_ERR_BAD_ARGUMENT, who);
        return false;
    }
    return true;
}

// operand -> result copy for in/out state when the caller did not alias the buffers
static void alias_copy(cudaStream_t s, void* result, const void* operand, size_t bytes) {
    if (result != operand && bytes) {
        const cudaError_t e = cudaMemcpyAsync(result, operand, bytes, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) set_error((int)e, "alias copy");
    }
}

static Grid make_grid(const VerosB200IsoDescriptor* d, void* dxt, void* dxu, void* dyt, void* dyu, void* cost,
                      void* cosu, void* dzt, void* dzw, void* zt) {
    Grid g;
    g.N = d->nx_tot;
    g.M = d->ny_tot;
    g.nz = d->nz;
    g.dxt = (const double*)dxt;
    g.dxu = (const double*)dxu;
    g.dyt = (const double*)dyt;
    g.dyu = (const double*)dyu;
    g.cost = (const double*)cost;
    g.cosu = (const double*)cosu;
    g.dzt = (const double*)dzt;
    g.dzw = (const double*)dzw;
    g.zt = (const double*)zt;
    return g;
}

cudaEvent_t* g_prof_events = nullptr;
int g_prof_n = 0;
void prof_mark(cudaStream_t s, int q) {
    if (g_prof_events && q < g_prof_n) cudaEventRecord(g_prof_events[q], s);
}

static int pre_variant(const VerosB200IsoDescriptor* d) {
    return (d->flags & VEROS_B200_FLAG_PRE_SINGLE) ? 1 : (d->flags & VEROS_B200_FLAG_PRE_SPLIT) ? 2 : 0;
}

static size_t tabs_doubles(const VerosB200IsoDescriptor* d) {
    return (tables_doubles(d->nx_tot, d->ny_tot, d->nz) + 1) & ~(size_t)1;  // keep 16 B alignment behind it
}
static size_t drd_doubles(const VerosB200IsoDescriptor* d) {
    return d->eq_of_state_type == 5 ? (size_t)2 * d->nx_tot * d->ny_tot * d->nz : 0;
}
static size_t pre_ws_doubles(const VerosB200IsoDescriptor* d) { return tabs_doubles(d) + drd_doubles(d); }
// fused step: contiguous copies of int_drhodT/S[..., tau] (PreArgs::stage)
static size_t step_stage_doubles(const VerosB200IsoDescriptor* d) {
    return d->enable_conserve_energy ? (size_t)2 * d->nx_tot * d->ny_tot * d->nz : 0;
}

}  // namespace vb

using namespace vb;

extern "C" {

void veros_b200_solve_implicit_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200SolveDescriptor>(opaque, len, "solve_implicit: bad descriptor");
    if (!d) return;
    if (d->num_systems < 0 || d->system_depth < 0) return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "solve_implicit: negative size");
    launch_solve_implicit((cudaStream_t)stream, d->num_systems, d->system_depth, (const double*)B[0], (const double*)B[1],
                          (const double*)B[2], (const double*)B[3], (const uint8_t*)B[4], (const uint8_t*)B[5],
                          (d->flags & VEROS_B200_HAS_B_EDGE) ? (const double*)B[6] : nullptr,
                          (d->flags & VEROS_B200_HAS_D_EDGE) ? (const double*)B[7] : nullptr, (double*)B[8]);
}

void veros_b200_tdma_zmajor_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200TridiagDescriptor>(opaque, len, "tdma_zmajor_f64: bad descriptor");
    if (!d) return;
    launch_tdma_zmajor_f64((cudaStream_t)stream, d->num_systems, d->system_depth, (const double*)B[0], (const double*)B[1],
                           (const double*)B[2], (const double*)B[3], (double*)B[4], (double*)B[5]);
}

void veros_b200_tdma_zmajor_f32(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200TridiagDescriptor>(opaque, len, "tdma_zmajor_f32: bad descriptor");
    if (!d) return;
    launch_tdma_zmajor_f32((cudaStream_t)stream, d->num_systems, d->system_depth, (const float*)B[0], (const float*)B[1],
                           (const float*)B[2], (const float*)B[3], (float*)B[4], (float*)B[5]);
}

void veros_b200_iso_pre_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200IsoDescriptor>(opaque, len, "iso_pre: bad descriptor");
    if (!valid_iso(d, "iso_pre: bad argument")) return;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n3 = (size_t)d->nx_tot * d->ny_tot * d->nz;
    for (int q = 0; q < 4; ++q) alias_copy(s, B[24 + q], B[17 + q], n3 * 4 * 8);
    for (int q = 4; q < 7; ++q) alias_copy(s, B[24 + q], B[17 + q], n3 * 8);
    PreArgs a;
    a.g = make_grid(d, B[8], B[9], B[10], B[11], B[12], B[13], B[14], B[15], B[16]);
    a.temp = (const double*)B[0];
    a.salt = (const double*)B[1];
    a.tau = (const int32_t*)B[2];
    a.K_iso = (const double*)B[3];
    a.maskT = (const uint8_t*)B[4];
    a.maskU = (const uint8_t*)B[5];
    a.maskV = (const uint8_t*)B[6];
    a.maskW = (const uint8_t*)B[7];
    a.Ai_ez = (double*)B[24];
    a.Ai_nz = (double*)B[25];
    a.Ai_bx = (double*)B[26];
    a.Ai_by = (double*)B[27];
    a.K_11 = (double*)B[28];
    a.K_22 = (double*)B[29];
    a.K_33 = (double*)B[30];
    a.tables = (double*)B[31];
    a.tables_ready = 0;
    a.dt_tracer = d->dt_tracer;
    a.drdT = a.tables + tabs_doubles(d);
    a.drdS = a.drdT + n3;
    a.with_flux = 0;
    a.with_stage = 0;
    a.no_mask_skip = (d->flags & VEROS_B200_FLAG_NO_MASK_SKIP) ? 1 : 0;
    a.variant = pre_variant(d);
    a.stage[0] = a.stage[1] = nullptr;
    a.stage_src[0] = a.stage_src[1] = nullptr;
    for (int t = 0; t < 2; ++t)
        for (int q = 0; q < 3; ++q) a.flux[t][q] = nullptr;
    a.eos = d->eq_of_state_type;
    a.K_iso_steep = d->K_iso_steep;
    a.iso_slopec = d->iso_slopec;
    a.iso_dslope = d->iso_dslope;
    launch_iso_pre(s, a);
}

void veros_b200_iso_diffusion_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200IsoDescriptor>(opaque, len, "iso_diffusion: bad descriptor");
    if (!valid_iso(d, "iso_diffusion: bad argument")) return;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n3 = (size_t)d->nx_tot * d->ny_tot * d->nz;
    const bool energy = d->enable_conserve_energy != 0;
    alias_copy(s, B[25], B[0], n3 * 3 * 8);
    alias_copy(s, B[26], B[1], n3 * 8);
    if (energy) alias_copy(s, B[27], B[2], n3 * 8);
    DiffArgs a;
    a.g = make_grid(d, B[16], B[17], B[18], B[19], B[20], B[21], B[22], B[23], nullptr);
    a.ntr = 1;
    a.t[0].tr = (double*)B[25];
    a.t[0].dtracer = (double*)B[26];
    a.t[0].int_drhodX = (const double*)B[24];
    a.t[1] = a.t[0];
    a.P_diss = (double*)B[27];
    a.tau = (const int32_t*)B[3];
    a.taup1 = (const int32_t*)B[4];
    a.K = (const double*)B[5];
    a.Ai_ez = (const double*)B[6];
    a.Ai_nz = (const double*)B[7];
    a.Ai_bx = (const double*)B[8];
    a.Ai_by = (const double*)B[9];
    a.K_11 = (const double*)B[10];
    a.K_22 = (const double*)B[11];
    a.K_33 = (const double*)B[12];
    a.maskT = (const uint8_t*)B[13];
    a.maskW = (const uint8_t*)B[14];
    a.kbot = (const int32_t*)B[15];
    a.skew = (d->flags & VEROS_B200_FLAG_SKEW) ? 1 : 0;
    a.energy = energy ? 1 : 0;
    a.skip_west_ring = (d->flags & VEROS_B200_FLAG_NO_WEST_RING) ? 1 : 0;
    a.skip_east_ring = (d->flags & VEROS_B200_FLAG_NO_EAST_RING) ? 1 : 0;
    a.fluxes_ready = 0;
    a.dry_skip = 0;  // stand-alone op: fluxes come from caller-supplied Ai_* / K_*, nothing is assumed about them
    a.stage_x[0] = a.stage_x[1] = nullptr;
    a.dt_tracer = d->dt_tracer;
    a.grav = d->grav;
    a.rho_0 = d->rho_0;
    double* ws = (double*)B[28];
    a.tables = ws + diffusion_workspace_doubles(d->nx_tot, d->ny_tot, d->nz, 1);
    launch_setup_tables(s, a.g, d->dt_tracer, a.tables);
    if (call_failed()) return;
    launch_iso_diffusion_ws(s, a, ws);
}

void veros_b200_iso_step_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200IsoDescriptor>(opaque, len, "iso_step: bad descriptor");
    if (!valid_iso(d, "iso_step: bad argument")) return;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n3 = (size_t)d->nx_tot * d->ny_tot * d->nz;
    const bool energy = d->enable_conserve_energy != 0;
    const size_t bytes[12] = {n3 * 24, n3 * 24, n3 * 8, n3 * 8, energy ? n3 * 8 : 0, n3 * 32, n3 * 32, n3 * 32, n3 * 32,
                              n3 * 8,  n3 * 8,  n3 * 8};
    for (int q = 0; q < 12; ++q) alias_copy(s, B[31 + q], B[q], bytes[q]);
    double* ws = (double*)B[43];

    PreArgs p;
    p.g = make_grid(d, B[20], B[21], B[22], B[23], B[24], B[25], B[26], B[27], B[28]);
    p.temp = (const double*)B[31];
    p.salt = (const double*)B[32];
    p.tau = (const int32_t*)B[12];
    p.K_iso = (const double*)B[14];
    p.maskT = (const uint8_t*)B[15];
    p.maskU = (const uint8_t*)B[16];
    p.maskV = (const uint8_t*)B[17];
    p.maskW = (const uint8_t*)B[18];
    p.Ai_ez = (double*)B[36];
    p.Ai_nz = (double*)B[37];
    p.Ai_bx = (double*)B[38];
    p.Ai_by = (double*)B[39];
    p.K_11 = (double*)B[40];
    p.K_22 = (double*)B[41];
    p.K_33 = (double*)B[42];
    // separate launches unless the fused persistent kernel is asked for (measured slower so far: DESIGN.md)
    const bool classic = !((d->flags & VEROS_B200_FLAG_STEP_FUSED) != 0 || getenv("VEROS_B200_STEP_FUSED") != nullptr);
    p.dt_tracer = d->dt_tracer;
    p.eos = d->eq_of_state_type;
    p.K_iso_steep = d->K_iso_steep;
    p.iso_slopec = d->iso_slopec;
    p.iso_dslope = d->iso_dslope;
    p.stage_src[0] = (const double*)B[29];
    p.stage_src[1] = (const double*)B[30];
    p.variant = pre_variant(d);
    p.no_mask_skip = (d->flags & VEROS_B200_FLAG_NO_MASK_SKIP) ? 1 : 0;

    DiffArgs a;
    a.g = p.g;
    a.ntr = 2;
    a.t[0].tr = (double*)B[31];
    a.t[0].dtracer = (double*)B[33];
    a.t[0].int_drhodX = (const double*)B[29];
    a.t[1].tr = (double*)B[32];
    a.t[1].dtracer = (double*)B[34];
    a.t[1].int_drhodX = (const double*)B[30];
    a.P_diss = (double*)B[35];
    a.tau = (const int32_t*)B[12];
    a.taup1 = (const int32_t*)B[13];
    a.K = (const double*)B[14];
    a.Ai_ez = p.Ai_ez;
    a.Ai_nz = p.Ai_nz;
    a.Ai_bx = p.Ai_bx;
    a.Ai_by = p.Ai_by;
    a.K_11 = p.K_11;
    a.K_22 = p.K_22;
    a.K_33 = p.K_33;
    a.maskT = p.maskT;
    a.maskW = p.maskW;
    a.kbot = (const int32_t*)B[19];
    a.skew = 0;
    a.energy = energy ? 1 : 0;
    a.skip_west_ring = (d->flags & VEROS_B200_FLAG_NO_WEST_RING) ? 1 : 0;
    a.skip_east_ring = (d->flags & VEROS_B200_FLAG_NO_EAST_RING) ? 1 : 0;
    a.dt_tracer = d->dt_tracer;
    a.grav = d->grav;
    a.rho_0 = d->rho_0;
    a.dry_skip = p.no_mask_skip ? 0 : 1;

    if (!classic) {
        // ---- the fused persistent kernel (iso_mega.cu): tables | queue + counters | scratch ring --------------------
        p.tables = ws;
        p.tables_ready = 1;
        unsigned int* sync = reinterpret_cast<unsigned int*>(ws + tabs_doubles(d));
        double* ring = ws + tabs_doubles(d) + mega_sync_doubles(d->nx_tot);
        a.tables = p.tables;
        a.fluxes_ready = 1;
        prof_mark(s, 0);
        launch_setup_tables(s, p.g, d->dt_tracer, p.tables, sync, (int)(2 * mega_sync_doubles(d->nx_tot)));
        if (call_failed()) return;
        prof_mark(s, 1);
        launch_iso_mega(s, p, a, ring, sync);
        prof_mark(s, 2);
        prof_mark(s, 3);
        return;
    }

    // ---- separate launches: setup, (TEOS-10), slope + flux kernel(s), update kernel --------------------------------
    double* stage = ws + diffusion_workspace_doubles(d->nx_tot, d->ny_tot, d->nz, 2);  // behind the flux/scratch arrays
    p.tables = stage + step_stage_doubles(d);
    p.tables_ready = 1;
    p.drdT = p.tables + tabs_doubles(d);
    p.drdS = p.drdT + n3;
    prof_mark(s, 0);
    launch_setup_tables(s, p.g, d->dt_tracer, p.tables);
    if (call_failed()) return;
    p.with_flux = 1;
    for (int t = 0; t < 2; ++t)
        for (int q = 0; q < 3; ++q) p.flux[t][q] = ws + (size_t)(3 * t + q) * n3;
    p.with_stage = energy ? 1 : 0;
    p.stage[0] = energy ? stage : nullptr;
    p.stage[1] = energy ? stage + n3 : nullptr;
    launch_iso_pre(s, p, /*profile=*/true);
    if (call_failed()) return;
    a.fluxes_ready = 1;
    a.stage_x[0] = p.stage[0];
    a.stage_x[1] = p.stage[1];
    a.tables = p.tables;
    launch_iso_diffusion_ws(s, a, ws);
    prof_mark(s, 3);
}

void veros_b200_vertmix_tempsalt_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200VmixDescriptor>(opaque, len, "vertmix_tempsalt: bad descriptor");
    if (!d) return;
    if (d->nx_tot < 0 || d->ny_tot < 0 || d->nz < 0 || !(d->dt_tracer > 0.0))
        return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "vertmix_tempsalt: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n3 = (size_t)d->nx_tot * d->ny_tot * d->nz;
    if (n3 == 0) return;
    alias_copy(s, B[9], B[0], n3 * 3 * 8);
    alias_copy(s, B[10], B[1], n3 * 3 * 8);
    VmixArgs a;
    a.N = d->nx_tot;
    a.M = d->ny_tot;
    a.nz = d->nz;
    a.temp = (double*)B[9];
    a.salt = (double*)B[10];
    a.taup1 = (const int32_t*)B[2];
    a.kappaH = (const double*)B[3];
    a.forc_temp = (const double*)B[4];
    a.forc_salt = (const double*)B[5];
    a.kbot = (const int32_t*)B[6];
    a.dzt = (const double*)B[7];
    a.dzw = (const double*)B[8];
    a.dtemp_vmix = (double*)B[11];
    a.dsalt_vmix = (double*)B[12];
    a.dt_tracer = d->dt_tracer;
    launch_vertmix(s, a);
}

void veros_b200_implicit_vert_friction_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200ColumnDescriptor>(opaque, len, "implicit_vert_friction: bad descriptor");
    if (!d) return;
    if (d->nx_tot < 4 || d->ny_tot < 4 || d->nz < 1 || !(d->dt > 0.0))
        return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "implicit_vert_friction: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n3 = (size_t)d->nx_tot * d->ny_tot * d->nz;
    alias_copy(s, B[17], B[0], n3 * 24);
    alias_copy(s, B[18], B[1], n3 * 24);
    for (int q = 2; q < 5; ++q) alias_copy(s, B[17 + q], B[q], n3 * 8);
    launch_implicit_vert_friction(s, d->nx_tot, d->ny_tot, d->nz, d->dt, B);
}

void veros_b200_iso_diag_streamfunction_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200ColumnDescriptor>(opaque, len, "iso_diag_streamfunction: bad descriptor");
    if (!d) return;
    if (d->nx_tot < 4 || d->ny_tot < 4 || d->nz < 1)
        return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "iso_diag_streamfunction: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n3 = (size_t)d->nx_tot * d->ny_tot * d->nz;
    alias_copy(s, B[5], B[3], n3 * 8);
    alias_copy(s, B[6], B[4], n3 * 8);
    launch_diag_streamfunction(s, d->nx_tot, d->ny_tot, d->nz, B);
}

void veros_b200_set_eke_diffusivities_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200EkeDescriptor>(opaque, len, "set_eke_diffusivities: bad descriptor");
    if (!d) return;
    if (d->nx_tot < 1 || d->ny_tot < 1 || d->nz < 1)
        return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "set_eke_diffusivities: bad argument");
    launch_set_eke_diffusivities((cudaStream_t)stream, d, B);
}

void veros_b200_advect_tempsalt_f64(void* stream, void** B, const char* opaque, size_t len) {
    begin_call();
    const auto* d = unpack<VerosB200AdvectDescriptor>(opaque, len, "advect_tempsalt: bad descriptor");
    if (!d) return;
    if (d->nx_tot < 5 || d->ny_tot < 5 || d->nz < 1 || !(d->dt_tracer > 0.0))
        return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "advect_tempsalt: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n3 = (size_t)d->nx_tot * d->ny_tot * d->nz;
    for (int q = 0; q < 4; ++q) alias_copy(s, B[19 + q], B[q], n3 * 24);
    launch_advect_tempsalt(s, d, B);
}

void veros_b200_profile_events(void** events, int n) {
    g_prof_events = reinterpret_cast<cudaEvent_t*>(events);
    g_prof_n = events ? n : 0;
}

size_t veros_b200_iso_pre_workspace_bytes(const char* opaque, size_t len) {
    const auto* d = unpack<VerosB200IsoDescriptor>(opaque, len, "iso_pre_workspace_bytes: bad descriptor");
    return d ? 8 * pre_ws_doubles(d) : 0;
}

size_t veros_b200_iso_diffusion_workspace_bytes(const char* opaque, size_t len) {
    const auto* d = unpack<VerosB200IsoDescriptor>(opaque, len, "iso_diffusion_workspace_bytes: bad descriptor");
    return d ? 8 * (diffusion_workspace_doubles(d->nx_tot, d->ny_tot, d->nz, 1) + tabs_doubles(d)) : 0;
}

size_t veros_b200_iso_step_workspace_bytes(const char* opaque, size_t len) {
    const auto* d = unpack<VerosB200IsoDescriptor>(opaque, len, "iso_step_workspace_bytes: bad descriptor");
    if (!d) return 0;
    // enough for either implementation of the step (the descriptor flags choose at call time)
    const size_t classic = pre_ws_doubles(d) + diffusion_workspace_doubles(d->nx_tot, d->ny_tot, d->nz, 2) + step_stage_doubles(d);
    const size_t fused = tabs_doubles(d) + mega_sync_doubles(d->nx_tot) +
                         mega_ring_doubles(d->nx_tot, d->ny_tot, d->nz, d->eq_of_state_type, d->enable_conserve_energy != 0);
    return 8 * (classic > fused ? classic : fused);
}

size_t veros_b200_iso_step_stats_offset(const char* opaque, size_t len) {
    const auto* d = unpack<VerosB200IsoDescriptor>(opaque, len, "iso_step_stats_offset: bad descriptor");
    return d ? 8 * (tabs_doubles(d) + mega_stats_offset_doubles(d->nx_tot)) : 0;
}

int veros_b200_last_error(void) { return g_err.load(); }
const char* veros_b200_last_error_string(void) { return g_err_text; }
void veros_b200_clear_error(void) {
    std::lock_guard<std::mutex> lock(g_err_mutex);
    g_err.store(0);
    g_err_text[0] = 0;
}
int veros_b200_abi_version(void) { return VEROS_B200_ABI_VERSION; }
unsigned long long veros_b200_launch_count(void) { return g_launches.load(); }

}  // extern "C"

// sizeof() of the descriptors as compiled into the library, for ABI checks from the host language
extern "C" size_t veros_b200_descriptor_size(int which) {
    switch (which) {
    case 0: return sizeof(VerosB200TridiagDescriptor);
    case 1: return sizeof(VerosB200SolveDescriptor);
    case 2: return sizeof(VerosB200IsoDescriptor);
    case 3: return sizeof(VerosB200VmixDescriptor);
    case 4: return sizeof(VerosB200ColumnDescriptor);
    case 5: return sizeof(VerosB200EkeDescriptor);
    case 6: return sizeof(VerosB200AdvectDescriptor);
    default: return 0;
    }
}
