// psim_force.cuh -- per-particle force accumulation with a canonical summation order.
#pragma once
#include "psim_device.cuh"

namespace psim {

// `visit(f, want_rank)` must call f(xj, yj, rank) once for every candidate neighbour in the
// particle's 3x3 cell neighbourhood (the particle itself may be among them; `rank` is the visit
// rank of the candidate's cell and is only inspected when want_rank is true).
//
// Fast path: contributions are summed in visiting order.  A sum of <= 2 terms starting from +0 is
// order independent, so only particles with >= 3 in-range neighbours take the second pass, which
// re-accumulates in ascending (cell visit rank, x, y) order -- the same order the oracle uses
// (oracle/psim_oracle.c, "Summation order").  Pairs at distance exactly 0 (the self pair the
// reference evaluates, serial.cpp:107, and coincident particles) contribute coef*0 = -0 and are
// skipped: a + (-0) == a.
template <class Visit>
__device__ __forceinline__ void accumulate_force(double xi, double yi, Visit&& visit, double& ax, double& ay,
                                                 int& neighbours) {
    double sx = 0.0, sy = 0.0;
    int cnt = 0;
    visit(
        [&](double xj, double yj, int) {
            const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi);
            const double r2 = pair_r2(dx, dy);
            if (r2 > kCutoff2 || r2 == 0.0) return;
            double cx, cy;
            pair_contrib(dx, dy, r2, cx, cy);
            sx = __dadd_rn(sx, cx);
            sy = __dadd_rn(sy, cy);
            ++cnt;
        },
        false);
    if (cnt >= 3) {
        sx = 0.0;
        sy = 0.0;
        NbKey last{-1, 0.0, 0.0};
        for (int guard = 0; guard < cnt; ++guard) {
            NbKey best{0, 0.0, 0.0};
            bool found = false;
            int mult = 0;
            double bx = 0.0, by = 0.0;
            visit(
                [&](double xj, double yj, int rank) {
                    const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi);
                    const double r2 = pair_r2(dx, dy);
                    if (r2 > kCutoff2 || r2 == 0.0) return;
                    const NbKey k{rank, xj, yj};
                    if (!key_less(last, k)) return;
                    if (!found || key_less(k, best)) {
                        best = k;
                        found = true;
                        mult = 1;
                        pair_contrib(dx, dy, r2, bx, by);
                    } else if (key_equal(k, best)) {
                        ++mult;
                    }
                },
                true);
            if (!found) break;
            for (int m = 0; m < mult; ++m) {
                sx = __dadd_rn(sx, bx);
                sy = __dadd_rn(sy, by);
            }
            last = best;
        }
    }
    ax = sx;
    ay = sy;
    neighbours = cnt;
}

}  // namespace psim
