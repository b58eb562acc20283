"""The strict kernels divide by grid metrics with  q = x*ry; r = fma(-q, y, x); q' = fma(r, ry, q)
(ry = correctly rounded 1/y, veros_b200/csrc/strict.cuh).  Check on the CPU, with the same operation
sequence in C, that this equals IEEE division on random and adversarial operands."""
import os
import subprocess
import tempfile

SRC = r"""
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
static uint64_t s = 88172645463325252ULL;
static int span = 30;  /* exponents are drawn from [-span, span] */
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double rd(int mode) {
    uint64_t m = rnd() & 0xFFFFFFFFFFFFFULL;
    if (mode) { int sh = rnd() % 52; if (rnd() & 1) m |= (~0ULL >> (12 + sh)) << sh; else m &= ~(~0ULL >> (12 + sh)); m &= 0xFFFFFFFFFFFFFULL; }
    int e = 1023 + (int)(rnd() % (2 * span + 1)) - span;
    uint64_t b = ((uint64_t)e << 52) | m; double d; *(uint64_t*)&d = b; return (rnd() & 1) ? d : -d;
}
int main(int argc, char** argv) {
    long bad = 0, n = 30000000L;
    int xspan = 30, yspan = 30;
    if (argc > 2) { xspan = atoi(argv[1]); yspan = atoi(argv[2]); }
    for (long i = 0; i < n; i++) {
        int mode = (i & 3) == 0;
        span = xspan; double x = rd(mode);
        span = yspan; double y = rd(mode);
        double ry = 1.0 / y, q = x * ry, r = fma(-q, y, x), q1 = fma(r, ry, q);
        if (q1 != x / y) bad++;
    }
    printf("%ld\n", bad);
    return 0;
}
"""


def test_markstein_division_is_correctly_rounded():
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "m.c"), os.path.join(d, "m")
        open(src, "w").write(SRC)
        subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-mfma", src, "-o", exe, "-lm"])
        assert subprocess.check_output([exe]).strip() == b"0"
        # the domain the column solve admits to its fast path (csrc/tdma_device.cuh): pivot exponent within
        # +-400, numerator exponent within +-500
        assert subprocess.check_output([exe, "500", "400"]).strip() == b"0"
