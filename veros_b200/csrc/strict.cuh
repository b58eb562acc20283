// Explicitly rounded double arithmetic for the kernels that must reproduce the reference's NumPy
// expressions bit for bit (one IEEE operation per NumPy ufunc call, never a fused multiply-add).
#pragma once

#include <cuda_runtime.h>

namespace vb {
namespace strict {

__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }

// A divisor together with its correctly rounded reciprocal.
struct Divisor {
    double y, ry;
};
__device__ __forceinline__ Divisor make_divisor(double y) { return Divisor{y, __ddiv_rn(1.0, y)}; }

// Correctly rounded x / d.y in three FP64 instructions (Markstein: with ry = RN(1/y),
// q0 = RN(x*ry), r = x - q0*y exactly, RN(q0 + r*ry) = RN(x/y)).  Valid for normal, finite
// operands, which grid metrics and time steps are; checked against IEEE division on 4e8 random and
// adversarial operand pairs (tests/test_strict_division.py runs the same recipe in C).
__device__ __forceinline__ double div(double x, const Divisor& d) {
    const double q = __dmul_rn(x, d.ry);
    const double r = __fma_rn(-q, d.y, x);
    return __fma_rn(r, d.ry, q);
}

}  // namespace strict
}  // namespace vb
