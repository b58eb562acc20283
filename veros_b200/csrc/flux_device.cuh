// Strict (bit-reproducible) evaluation of the Griffies triad fluxes, shared by the fused
// slope+flux kernel (iso_pre.cu) and the stand-alone flux kernel (iso_diffusion.cu).
// Operation order follows veros/core/isoneutral/diffusion.py:25-111 term by term.
#pragma once

#include "strict.cuh"

namespace vb {

// East or north face (diffusion.py:25-47 / 52-77) without the cosu factor of the north flux:
//   sumz = sum_{kr} sum_{ip} diffloc * Ai[ip][kr] * (tr[ip][k+kr] - tr[ip][k+kr-1])
//   flux = sumz / (4 dzt) + (tr[1][k] - tr[0][k]) / metric * K
// A0 = Ai[ip=0][kr=0,1], A1 = Ai[ip=1][kr=0,1]; dz0_x / dz1_x are the lower / upper vertical
// differences in column ip = x (already clamped at the bottom and the surface, pad_z_edges).
__device__ __forceinline__ double flux_face(double diffloc, double A00, double A01, double A10, double A11,
                                            double dz0_0, double dz1_0, double dz0_1, double dz1_1, double dh,
                                            const strict::Divisor& d4zt, const strict::Divisor& dmetric,
                                            double Kxx) {
    using namespace strict;
    double sumz = add(0.0, mul(mul(diffloc, A00), dz0_0));
    sumz = add(sumz, mul(mul(diffloc, A10), dz0_1));
    sumz = add(sumz, mul(mul(diffloc, A01), dz1_0));
    sumz = add(sumz, mul(mul(diffloc, A11), dz1_1));
    return add(strict::div(sumz, d4zt), mul(strict::div(dh, dmetric), Kxx));
}

// Top face (diffusion.py:85-111): K31 / K32 part of the vertical flux.
//   dx[ip][kr] = tr(i+ip, j, k+kr) - tr(i-1+ip, j, k+kr),  dy[jp][kr] likewise in y.
__device__ __forceinline__ double flux_top(double diffloc, const double X[2][2], const double Y[2][2],
                                           const double dx[2][2], const double dy[2][2], double cosu0, double cosu1,
                                           const strict::Divisor& dcost, const strict::Divisor& d4xt,
                                           const strict::Divisor& d4ytc) {
    using namespace strict;
    double sumx = add(0.0, mul(strict::div(mul(diffloc, X[0][0]), dcost), dx[0][0]));
    sumx = add(sumx, mul(strict::div(mul(diffloc, X[0][1]), dcost), dx[0][1]));
    sumx = add(sumx, mul(strict::div(mul(diffloc, X[1][0]), dcost), dx[1][0]));
    sumx = add(sumx, mul(strict::div(mul(diffloc, X[1][1]), dcost), dx[1][1]));
    double sumy = add(0.0, mul(mul(mul(diffloc, Y[0][0]), cosu0), dy[0][0]));
    sumy = add(sumy, mul(mul(mul(diffloc, Y[0][1]), cosu0), dy[0][1]));
    sumy = add(sumy, mul(mul(mul(diffloc, Y[1][0]), cosu1), dy[1][0]));
    sumy = add(sumy, mul(mul(mul(diffloc, Y[1][1]), cosu1), dy[1][1]));
    return add(strict::div(sumx, d4xt), strict::div(sumy, d4ytc));
}

}  // namespace vb
