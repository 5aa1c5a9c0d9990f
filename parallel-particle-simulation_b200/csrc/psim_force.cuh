// psim_force.cuh -- per-particle force accumulation with a canonical summation order.
#pragma once
#include "psim_device.cuh"

namespace psim {

constexpr int kMaxCanonical = 8;  // in-range neighbours sorted in local memory (beyond: selection sweeps)

// Rare path (>= 3 in-range neighbours): order the contributions by (cell visit rank, x, y) and sum.
// Out of line so that its FP64 division / square-root sequences exist once in the kernel image.
static __device__ __noinline__ void canonical_sum(int n, const int* rank, const double* dx, const double* dy, const double* r2,
                                           const double* xj, const double* yj, double* out_ax, double* out_ay) {
    int order[kMaxCanonical];
    for (int a = 0; a < n; ++a) {
        int p = a;
        while (p > 0) {
            const int b = order[p - 1];
            const NbKey ka{rank[a], xj[a], yj[a]}, kb{rank[b], xj[b], yj[b]};
            if (!key_less(ka, kb)) break;
            order[p] = b;
            --p;
        }
        order[p] = a;
    }
    double sx = 0.0, sy = 0.0;
    for (int a = 0; a < n; ++a) {
        const int k = order[a];
        double cx, cy;
        pair_contrib(dx[k], dy[k], r2[k], cx, cy);
        sx = __dadd_rn(sx, cx);
        sy = __dadd_rn(sy, cy);
    }
    *out_ax = sx;
    *out_ay = sy;
}

// out-of-line single contribution (only for the > kMaxCanonical overflow tail)
static __device__ __noinline__ void pair_contrib_outlined(double dx, double dy, double r2, double* cx, double* cy) {
    pair_contrib(dx, dy, r2, *cx, *cy);
}

// `visit(f)` must call f(xj, yj, tag) once for every candidate neighbour in the particle's 3x3
// cell neighbourhood (the particle itself may be among them); `rank_of(xj, yj, tag)` returns the
// visit rank of that candidate's cell in the reference's order self,T,B,L,R,TL,TR,BL,BR and is only
// evaluated on the rare canonical path.
//
// Pass 1 only measures distances and remembers the first two in-range neighbours; the expensive
// coefficient (sqrt + three divisions, reference serial.cpp:29-33) is then evaluated at ONE code
// site.  A sum of <= 2 terms starting from +0 is order independent, so only particles with >= 3
// in-range neighbours take the canonical path, which re-collects and sums in ascending
// (cell visit rank, x, y) order -- the order the oracle uses (oracle/psim_oracle.c, "Summation
// order").  Pairs at distance exactly 0 (the self pair the reference evaluates, serial.cpp:107, and
// coincident particles) contribute coef*0 = -0 and are skipped: a + (-0) == a.
template <class Visit, class RankOf>
__device__ __forceinline__ void accumulate_force(double xi, double yi, Visit&& visit, RankOf&& rank_of, double& ax,
                                                 double& ay, int& neighbours) {
    int cnt = 0;
    double dx0 = 0.0, dy0 = 0.0, r20 = 0.0, dx1 = 0.0, dy1 = 0.0, r21 = 0.0;
    visit([&](double xj, double yj, int) {
        const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi);
        const double r2 = pair_r2(dx, dy);
        if (r2 > kCutoff2 || r2 == 0.0) return;
        if (cnt == 0) {
            dx0 = dx; dy0 = dy; r20 = r2;
        } else if (cnt == 1) {
            dx1 = dx; dy1 = dy; r21 = r2;
        }
        ++cnt;
    });
    double sx = 0.0, sy = 0.0;
    if (cnt > kMaxCanonical) {
        // more in-range neighbours than any physical configuration has (dense synthetic inputs): selection
        // sort by repeated sweeps -- O(cnt^2) but exact for any count
        NbKey last{-1, 0.0, 0.0};
        for (int guard = 0; guard < cnt; ++guard) {
            NbKey best{0, 0.0, 0.0};
            bool found = false;
            int mult = 0;
            double bdx = 0.0, bdy = 0.0, br2 = 0.0;
            visit([&](double xj, double yj, int tag) {
                const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi);
                const double r2 = pair_r2(dx, dy);
                if (r2 > kCutoff2 || r2 == 0.0) return;
                const NbKey k{rank_of(xj, yj, tag), xj, yj};
                if (!key_less(last, k)) return;
                if (!found || key_less(k, best)) {
                    best = k; found = true; mult = 1; bdx = dx; bdy = dy; br2 = r2;
                } else if (key_equal(k, best)) {
                    ++mult;
                }
            });
            if (!found) break;
            double cx, cy;
            pair_contrib_outlined(bdx, bdy, br2, &cx, &cy);
            for (int m = 0; m < mult; ++m) {
                sx = __dadd_rn(sx, cx);
                sy = __dadd_rn(sy, cy);
            }
            last = best;
        }
    } else if (cnt >= 3) {
        int rank[kMaxCanonical];
        double dxs[kMaxCanonical], dys[kMaxCanonical], r2s[kMaxCanonical], xs[kMaxCanonical], ys[kMaxCanonical];
        int m = 0;
        visit([&](double xj, double yj, int tag) {
            const double dx = __dsub_rn(xj, xi), dy = __dsub_rn(yj, yi);
            const double r2 = pair_r2(dx, dy);
            if (r2 > kCutoff2 || r2 == 0.0) return;
            if (m < kMaxCanonical) {
                rank[m] = rank_of(xj, yj, tag); dxs[m] = dx; dys[m] = dy; r2s[m] = r2; xs[m] = xj; ys[m] = yj;
                ++m;
            }
        });
        canonical_sum(m, rank, dxs, dys, r2s, xs, ys, &sx, &sy);
    } else {
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) {
            double cx, cy;
            pair_contrib(k == 0 ? dx0 : dx1, k == 0 ? dy0 : dy1, k == 0 ? r20 : r21, cx, cy);
            sx = __dadd_rn(sx, cx);
            sy = __dadd_rn(sy, cy);
        }
    }
    ax = sx;
    ay = sy;
    neighbours = cnt;
}

}  // namespace psim
