// Packing of the 2-cell west/east halos of the tracers for the NCCL exchange (host-side plumbing of the
// multi-GPU harness, veros_b200/decomp.py).  The reference exchanges one time level of the (N,M,nz,3)
// tracers (veros/core/thermodynamics.py:293-298 -> veros/distributed.py:218-326); that level is strided
// in memory (24 bytes), so the planes are gathered into one contiguous send buffer per direction and
// scattered from one receive buffer per direction: two launches per step for any number of fields.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace vb {

struct HaloFields {
    double* f[4];
    int n;
};

// mode 0: pack   west_buf <- planes [2,4),   east_buf <- planes [N-4,N-2)
// mode 1: unpack planes [0,2) <- west_buf,   planes [N-2,N) <- east_buf
template <int MODE>
__global__ void __launch_bounds__(256)
halo_kernel(HaloFields h, int N, size_t plane, int nlev, int level, double* __restrict__ west, double* __restrict__ east,
            int do_west, int do_east) {
    const size_t per_field = 2 * plane;
    const size_t total = per_field * h.n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int fi = (int)(e / per_field);
        const size_t r = e - fi * per_field;  // index inside the two planes
        double* f = h.f[fi];
        const size_t w_cell = (MODE == 0 ? 2 * plane : 0) + r;
        const size_t e_cell = (MODE == 0 ? (size_t)(N - 4) * plane : (size_t)(N - 2) * plane) + r;
        if (MODE == 0) {
            if (do_west) west[e] = f[w_cell * nlev + level];
            if (do_east) east[e] = f[e_cell * nlev + level];
        } else {
            if (do_west) f[w_cell * nlev + level] = west[e];
            if (do_east) f[e_cell * nlev + level] = east[e];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Halo exchange over peer memory (NVLink / NVSwitch): no staging buffers, no NCCL call.  Every rank runs
// this kernel on its own stream; the neighbours' arrays and flag words are mapped through CUDA IPC.
//   1. tell both neighbours "my ghost planes may be overwritten for exchange number `seq`" (READY)
//   2. wait for their READY, then store my edge planes [2,4) / [N-4,N-2) straight into the east ghosts of the
//      west neighbour / the west ghosts of the east neighbour
//   3. the last CTA to finish fences system-wide, tells both neighbours DONE and waits for their DONE, so when
//      the kernel completes this rank's ghost planes hold the neighbours' data of exchange `seq`.
// Flags are monotonically increasing exchange numbers, so nothing is ever reset.  No CTA waits for another
// CTA of its own grid, so the grid need not be co-resident.
// flags layout (int32): [0] READY from west, [1] READY from east, [2] DONE from west, [3] DONE from east.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void flag_store(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_load(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Bounded wait: ranks that disagree about the number of exchanges (a host-side bug) must produce an error, not a GPU
// that spins until the job is killed.
__device__ __forceinline__ void flag_wait(const int* p, int seq) {
    if (flag_load(p) >= seq) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    unsigned spins = 0;
    while (flag_load(p) < seq) {
        __nanosleep(100);
        if ((++spins & 0x3ffu) == 0u) {
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > 30000000000ull) __trap();  // 30 s
        }
    }
}

struct PeerHalo {
    double* mine[4];
    double* west[4];  // the west neighbour's arrays (null: no neighbour on that side)
    double* east[4];
    int n;
};

__global__ void __launch_bounds__(256)
halo_put_kernel(PeerHalo h, int N, int N_west, size_t plane, int nlev, int level, int seq, int* my_flags,
                int* west_flags, int* east_flags, unsigned int* counter) {
    const bool has_w = west_flags != nullptr, has_e = east_flags != nullptr;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            if (has_w) flag_store(west_flags + 1, seq);  // I am their east neighbour
            if (has_e) flag_store(east_flags + 0, seq);
        }
        if (has_w) flag_wait(my_flags + 0, seq);
        if (has_e) flag_wait(my_flags + 1, seq);
    }
    __syncthreads();
    const size_t per_field = 2 * plane;
    const size_t total = per_field * h.n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int fi = (int)(e / per_field);
        const size_t r = e - fi * per_field;
        const double* f = h.mine[fi];
        if (has_w) h.west[fi][((size_t)(N_west - 2) * plane + r) * nlev + level] = f[(2 * plane + r) * nlev + level];
        if (has_e) h.east[fi][r * nlev + level] = f[((size_t)(N - 4) * plane + r) * nlev + level];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {  // every CTA's stores are fenced: publish, then wait for the neighbours
            *counter = 0u;
            __threadfence_system();
            if (has_w) flag_store(west_flags + 3, seq);
            if (has_e) flag_store(east_flags + 2, seq);
            if (has_w) flag_wait(my_flags + 2, seq);
            if (has_e) flag_wait(my_flags + 3, seq);
        }
    }
}

}  // namespace vb

using namespace vb;

// CUDA IPC plumbing for the peer-memory exchange.  The importing side must open the handle with ITS compute
// device current, so that the mapping (and the lazily enabled peer access) belongs to the context its kernels
// run in; a mapping made in the exporter's device context of the importing process is not reachable from there.
extern "C" int veros_b200_ipc_get_handle(void* base, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    const cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), base);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error((int)e, "ipc_get_handle");
        return 1;
    }
    return 0;
}

extern "C" void* veros_b200_ipc_open_handle(int device, const void* handle64) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error((int)e, "ipc_open_handle");
        return nullptr;
    }
    return p;
}

extern "C" void veros_b200_ipc_close(void* base) {
    if (base) cudaIpcCloseMemHandle(base);
}

extern "C" void veros_b200_halo_put(void* stream, int seq, void** fields, void** west_fields, void** east_fields,
                                    int nfields, int N, int N_west, int M, int nz, int nlev, int level, void* my_flags,
                                    void* west_flags, void* east_flags, void* counter) {
    if (nfields < 1 || nfields > 4 || N < 8 || nlev < 1 || level < 0 || level >= nlev || seq < 1 || !my_flags || !counter ||
        (west_flags && (!west_fields || N_west < 8)) || (east_flags && !east_fields))
        return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "halo_put: bad argument");
    PeerHalo h;
    h.n = nfields;
    for (int q = 0; q < 4; ++q) {
        h.mine[q] = q < nfields ? (double*)fields[q] : nullptr;
        h.west[q] = (q < nfields && west_flags) ? (double*)west_fields[q] : nullptr;
        h.east[q] = (q < nfields && east_flags) ? (double*)east_fields[q] : nullptr;
    }
    const size_t plane = (size_t)M * nz;
    const size_t total = 2 * plane * nfields;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, 296));
    halo_put_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h, N, N_west, plane, nlev, level, seq, (int*)my_flags,
                                                           (int*)west_flags, (int*)east_flags, (unsigned int*)counter);
    count_launch();
    check_launch("halo_put_kernel");
}

extern "C" void veros_b200_halo_pack_unpack(void* stream, int mode, void** fields, int nfields, int N, int M, int nz,
                                            int nlev, int level, void* west_buf, void* east_buf) {
    if (nfields < 1 || nfields > 4 || N < 8 || nlev < 1 || level < 0 || level >= nlev)
        return set_error(VEROS_B200_ERR_BAD_ARGUMENT, "halo_pack_unpack: bad argument");
    HaloFields h;
    h.n = nfields;
    for (int q = 0; q < 4; ++q) h.f[q] = q < nfields ? (double*)fields[q] : nullptr;
    const size_t plane = (size_t)M * nz;
    const size_t total = 2 * plane * nfields;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == 0)
        halo_kernel<0><<<grid, 256, 0, s>>>(h, N, plane, nlev, level, (double*)west_buf, (double*)east_buf,
                                            west_buf != nullptr, east_buf != nullptr);
    else
        halo_kernel<1><<<grid, 256, 0, s>>>(h, N, plane, nlev, level, (double*)west_buf, (double*)east_buf,
                                            west_buf != nullptr, east_buf != nullptr);
    count_launch();
    check_launch("halo_kernel");
}
