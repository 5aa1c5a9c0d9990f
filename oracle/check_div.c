/*
 * check_div.c -- TEST INFRASTRUCTURE: pins the division-free cell index of the CUDA kernels.
 *
 * The reference computes a cell coordinate as floor(x / 0.01) with an IEEE double division
 * (reference part1/serial.cpp:41-42).  The kernels (csrc/psim_device.cuh: div_by_bin) evaluate
 *     q0 = RN(x * 100);  r = fma(-q0, 0.01, x);  q = fma(r, 100, q0)
 * and claim q == RN(x / 0.01) bit for bit (Markstein's theorem: q0 is a faithful quotient, the residual is
 * exact in one FMA, and 100.0 == RN(1 / 0.01)).  This program compares the two on
 *   1. every double within 200 ulps of every cell edge k * 0.01, k = 0 .. 40000  (16 M values),
 *   2. N random positions in [0, 400)                                              (argv[1], default 2e9),
 *   3. N/4 magnitudes spread log-uniformly over 1e-12 .. 2e12 with both signs (the force coefficient is divided by
 *      mass == 0.01 as well),
 *   4. a few special values,
 * and exits non-zero on any mismatch other than the sign of a zero quotient.
 *
 * Build: gcc -O2 -ffp-contract=off -o check_div check_div.c -lm   (fma() from libm is exact with or without hardware FMA)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline double div_by_bin(double x) {
    const double d = 0.01;
    const double q0 = x * 100.0;
    const double r = fma(-q0, d, x);
    return fma(r, 100.0, q0);
}
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

static long check(double x, long *bad) {
    const double a = x / 0.01, b = div_by_bin(x);
    if (memcmp(&a, &b, 8) != 0 && !(a == 0.0 && b == 0.0)) {
        if (*bad < 10) printf("MISMATCH x=%.17g div=%.17g fma=%.17g\n", x, a, b);
        ++*bad;
    }
    return 1;
}

int main(int argc, char **argv) {
    const long nrand = argc > 1 ? atol(argv[1]) : 2000000000L;
    long bad = 0, n = 0;
    for (int k = 0; k <= 40000; ++k) {
        double x = k * 0.01;
        for (int j = 0; j < 200; ++j) x = nextafter(x, -1.0);
        if (x < 0) x = 0;
        for (int j = 0; j < 400; ++j) { n += check(x, &bad); x = nextafter(x, 1e9); }
    }
    for (long i = 0; i < nrand; ++i) n += check((double)(rnd() >> 11) * (1.0 / 9007199254740992.0) * 400.0, &bad);
    /* 4. the force coefficient is also divided by mass == 0.01 (reference part1/serial.cpp:33): magnitudes up to 1e10, both signs */
    for (long i = 0; i < nrand / 4; ++i) {
        const double u = (double)(rnd() >> 11) * (1.0 / 9007199254740992.0);
        const double v = (double)(rnd() >> 11) * (1.0 / 9007199254740992.0);
        const double x = (1.0 + v) * pow(10.0, -12.0 + 24.0 * u);
        n += check(x, &bad);
        n += check(-x, &bad);
    }
    const double sp[] = {0.0, -0.0, 1e-300, 1e-310, 4.9e-324, 1e-20, 0.01, 0.02, 100.0, 282.84271247461902, 399.99999999999994};
    for (unsigned i = 0; i < sizeof sp / sizeof *sp; ++i) n += check(sp[i], &bad);
    printf("checked %ld values, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
