// Packing of the 2-cell west/east halos of the tracers for the NCCL exchange (host-side plumbing of the
// multi-GPU harness, veros_b200/decomp.py).  The reference exchanges one time level of the (N,M,nz,3)
// tracers (veros/core/thermodynamics.py:293-298 -> veros/distributed.py:218-326); that level is strided
// in memory (24 bytes), so the planes are gathered into one contiguous send buffer per direction and
// scattered from one receive buffer per direction: two launches per step for any number of fields.
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

}  // namespace vb

using namespace vb;

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
