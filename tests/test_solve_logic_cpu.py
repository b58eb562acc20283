"""The shipped column solve (veros_b200/csrc/tdma_device.cuh) compiled for the HOST -- the CUDA intrinsics
replaced by their IEEE definitions, the reciprocal by 1/y -- against a verbatim dgtsv on a few hundred thousand
random columns sprinkled with signed zeros, subnormals, huge values, infinities, NaNs and weak diagonals
(interchanges).  This exercises the control logic of the fast path (which levels are ordinary, where the
reciprocals travel, when the general loop takes over) at a scale the GPU tests do not, without a GPU.
The GPU tests check the same source with the real intrinsics (test_gpu_parity.py)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#define VB_HOST_EMULATION 1
#define __device__
#define __forceinline__ inline
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
using std::fabs;
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) {
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
}
#include "tdma_device.cuh"

// LAPACK dgtsv, NRHS right-hand sides, rows [k0, n) (dgtsv.f; same statement order as oracle/iso_oracle.c)
static void dgtsv_plain(int k0, int n, int nrhs, double* dl, double* d, double* du, double* b0, double* b1) {
    double* b[2] = {b0, b1};
    for (int k = k0; k < n - 1; k++) {
        if (fabs(d[k]) >= fabs(dl[k])) {
            double fact = dl[k] / d[k];
            d[k + 1] = d[k + 1] - fact * du[k];
            for (int j = 0; j < nrhs; j++) b[j][k + 1] = b[j][k + 1] - fact * b[j][k];
            dl[k] = 0.0;
        } else {
            double fact = d[k] / dl[k];
            d[k] = dl[k];
            double temp = d[k + 1];
            d[k + 1] = du[k] - fact * temp;
            dl[k] = du[k + 1];
            du[k + 1] = -fact * dl[k];
            du[k] = temp;
            for (int j = 0; j < nrhs; j++) {
                temp = b[j][k];
                b[j][k] = b[j][k + 1];
                b[j][k + 1] = temp - fact * b[j][k + 1];
            }
        }
    }
    for (int j = 0; j < nrhs; j++) {
        b[j][n - 1] = b[j][n - 1] / d[n - 1];
        if (n - 2 >= k0) b[j][n - 2] = (b[j][n - 2] - du[n - 2] * b[j][n - 1]) / d[n - 2];
        for (int k = n - 3; k >= k0; k--) b[j][k] = (b[j][k] - du[k] * b[j][k + 1] - dl[k] * b[j][k + 2]) / d[k];
    }
}

static uint64_t s = 0x9E3779B97F4A7C15ULL;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double uni() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }
static const double specials[] = {0.0, -0.0, 5e-324, -3e-310, 1e-300, -1e300, 1e308, INFINITY, -INFINITY, NAN,
                                  1e-170, 7e200, 1e-130, -2e150};
static inline double maybe_special(double v, double p) {
    return uni() < p ? specials[rnd() % (sizeof(specials) / sizeof(double))] : v;
}
static bool same(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::isnan(a) && std::isnan(b);
    return memcmp(&a, &b, 8) == 0;  // value and sign of zero
}

int main() {
    long bad = 0, fast_levels = 0, columns = 300000;
    for (long c = 0; c < columns; c++) {
        const int n = 1 + (int)(rnd() % 40), k0 = (int)(rnd() % n), nrhs = 1 + (int)(rnd() & 1);
        const double p = (c % 3 == 0) ? 0.0 : 0.03;             // a third of the columns stay clean
        const double weak = (c % 5 == 0) ? 1e-3 : 1.0;          // weak diagonals: interchanges
        const int pad = 2;
        std::vector<double> L(n + 2 * pad), D(n + 2 * pad), U(n + 2 * pad), R0(n + 2 * pad), R1(n + 2 * pad);
        for (int k = 0; k < n; k++) {
            const double a = -(0.1 + uni()), cc = -(0.1 + uni());
            L[pad + k] = maybe_special((rnd() % 7 == 0) ? 0.0 : a, p);     // zero couplings are common (K_33 = 0)
            U[pad + k] = (k == n - 1) ? 0.0 : maybe_special(cc, p);
            D[pad + k] = maybe_special((1.0 - a - cc) * ((rnd() & 3) ? 1.0 : weak), p);
            R0[pad + k] = maybe_special(10.0 * (uni() - 0.5), p);
            R1[pad + k] = maybe_special(35.0 + uni(), p);
        }
        std::vector<double> l2 = L, d2 = D, u2 = U, r02 = R0, r12 = R1;
        dgtsv_plain(k0, n, nrhs, l2.data() + pad, d2.data() + pad, u2.data() + pad, r02.data() + pad, r12.data() + pad);
        if (nrhs == 1)
            vb::dgtsv_column<1>(k0, n, L.data() + pad, D.data() + pad, U.data() + pad, R0.data() + pad, nullptr);
        else
            vb::dgtsv_column<2>(k0, n, L.data() + pad, D.data() + pad, U.data() + pad, R0.data() + pad, R1.data() + pad);
        for (int k = k0; k < n; k++) {
            if (!same(R0[pad + k], r02[pad + k])) bad++;
            if (nrhs == 2 && !same(R1[pad + k], r12[pad + k])) bad++;
            // a row that went through the fast path carries its pivot's reciprocal in the L slot
            if (k < n - 1 && L[pad + k] != 0.0 && same(L[pad + k], 1.0 / D[pad + k])) fast_levels++;
        }
    }
    printf("%ld %ld\n", bad, fast_levels);
    return 0;
}
"""


def test_shipped_column_solve_on_the_host_equals_dgtsv():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "cuda_runtime.h"), "w").write("// host stand-in\n")
        src, exe = os.path.join(d, "h.cpp"), os.path.join(d, "h")
        open(src, "w").write(HARNESS)
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-I", d,
                               "-I", os.path.join(ROOT, "veros_b200", "csrc"), src, "-o", exe])
        bad, fast = map(int, subprocess.check_output([exe]).split())
    assert bad == 0
    assert fast > 1_000_000  # the fast path was what ran on the ordinary levels
