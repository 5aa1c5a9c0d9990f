// psim_device.cuh -- device-side arithmetic shared by every kernel.
//
// All floating-point work is IEEE double with each product and sum rounded separately
// (__dmul_rn/__dadd_rn are never contracted into DFMA), because the x86-64 reference build does
// not contract either and the system is chaotic: enabling FMA on the same source decorrelates the
// 1000-step trajectory (Appendix B).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/psim_common.h"

namespace psim {

constexpr double kCutoff  = PSIM_CUTOFF;
constexpr double kCutoff2 = PSIM_CUTOFF * PSIM_CUTOFF;     // reference part1/serial.cpp:26
constexpr double kMinR2   = PSIM_MIN_R * PSIM_MIN_R;       // reference part1/serial.cpp:29
constexpr double kMass    = PSIM_MASS;
constexpr double kDt      = PSIM_DT;
constexpr double kBin     = PSIM_BIN_SIZE;

// Cell coordinate along one axis: floor(v / 0.01) with a correctly rounded IEEE division
// (reference part1/serial.cpp:41-42; v*100.0 is NOT bit-identical).
//
// div_by_bin(v) == RN(v / 0.01) bit for bit, without a division: q0 = RN(v * 100) is a faithful quotient
// (0.01 is 2.1e-17 relative above 1/100, plus one rounding: < 0.7 ulp), r = v - q0 * 0.01 is exact in one FMA,
// and 100.0 == RN(1 / 0.01), so RN(q0 + r * 100) is the correctly rounded quotient (Markstein's theorem).
// oracle/check_div.c compares it with the hardware division on every double within 200 ulps of a cell edge
// up to cell 40 000 and on 2e9 random positions: no mismatch (only the sign of a zero result differs).
// Two FMAs instead of a ~20-instruction division with a slow-path call, and no data-dependent branch.
__device__ __forceinline__ double div_by_bin(double v) {
    const double q0 = __dmul_rn(v, 100.0);
    const double r = __fma_rn(-q0, kBin, v);
    return __fma_rn(r, 100.0, q0);
}

// Clamped to [0, bincnt-1]: the reference indexes out of bounds for v == size when size/0.01 is an integer
// (e.g. 20 M particles, size == 100.0); that state is unreachable in practice, the clamp only keeps memory safe.
__device__ __forceinline__ int axis_cell(double v, int bincnt) {
    return min(max(__double2int_rd(div_by_bin(v)), 0), bincnt - 1);
}

__device__ __forceinline__ void cell_of(double x, double y, int bincnt, int& row, int& col) {
    row = axis_cell(x, bincnt);
    col = axis_cell(y, bincnt);
}

// Rank of a neighbour cell (dr, dc) in the reference's visiting order
// self, T, B, L, R, TL, TR, BL, BR (reference part1/serial.cpp:107-115; T = row-1, L = col-1).
__device__ __forceinline__ int visit_rank(int dr, int dc) {
    // index (dr+1)*3 + (dc+1):  (-1,-1)=TL5 (-1,0)=T1 (-1,1)=TR6 (0,-1)=L3 (0,0)=0 (0,1)=R4 (1,-1)=BL7 (1,0)=B2 (1,1)=BR8
    const unsigned long long table = 0x827403615ull;  // nibbles, lowest first: 5,1,6,3,0,4,7,2,8
    return (int)((table >> (4 * ((dr + 1) * 3 + (dc + 1)))) & 0xF);
}

// r2 of the pair and the in-range test (reference part1/serial.cpp:21-27).
__device__ __forceinline__ double pair_r2(double dx, double dy) {
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// Contribution of an in-range neighbour at offset (dx, dy): coef * d (reference serial.cpp:29-35).
__device__ __forceinline__ void pair_contrib(double dx, double dy, double r2, double& cx, double& cy) {
    r2 = fmax(r2, kMinR2);
    const double r = __dsqrt_rn(r2);
    // the last division is by mass == 0.01: the same exact two-FMA quotient as the cell index (oracle/check_div.c also
    // sweeps the coefficient's range, both signs); a zero quotient may differ in sign only, which a sum from +0 absorbs
    static_assert(PSIM_MASS == PSIM_BIN_SIZE, "div_by_bin divides by 0.01");
    const double coef = div_by_bin(__ddiv_rn(__dsub_rn(1.0, __ddiv_rn(kCutoff, r)), r2));
    cx = __dmul_rn(coef, dx);
    cy = __dmul_rn(coef, dy);
}

// Bounce a particle off the walls (reference part1/serial.cpp:53-61).
__device__ __forceinline__ void reflect_particle(double& x, double& y, double& vx, double& vy, double size) {
    if (x < 0 || x > size || y < 0 || y > size) {   // one rare branch on the common path
        const double two_size = __dmul_rn(2.0, size);
        while (x < 0 || x > size) {
            x = x < 0 ? -x : __dsub_rn(two_size, x);
            vx = -vx;
        }
        while (y < 0 || y > size) {
            y = y < 0 ? -y : __dsub_rn(two_size, y);
            vy = -vy;
        }
    }
}

// Integrate one particle and bounce it off the walls (reference part1/serial.cpp:46-61).
__device__ __forceinline__ void move_particle(double& x, double& y, double& vx, double& vy, double ax, double ay,
                                              double size) {
    vx = __dadd_rn(vx, __dmul_rn(ax, kDt));
    vy = __dadd_rn(vy, __dmul_rn(ay, kDt));
    x = __dadd_rn(x, __dmul_rn(vx, kDt));
    y = __dadd_rn(y, __dmul_rn(vy, kDt));
    reflect_particle(x, y, vx, vy, size);
}

// Canonical neighbour key: (visit rank of the neighbour's cell, neighbour x, neighbour y).
// Sums of >= 3 contributions are taken in ascending key order so that the result is a pure
// function of the particle set -- independent of storage order, tile size and slab count.
// (The reference's own in-cell order is hash-set iteration order and not reproducible.)
struct NbKey {
    int rank;
    double x, y;
};
__device__ __forceinline__ bool key_less(const NbKey& a, const NbKey& b) {
    if (a.rank != b.rank) return a.rank < b.rank;
    if (a.x != b.x) return a.x < b.x;
    return a.y < b.y;
}
__device__ __forceinline__ bool key_equal(const NbKey& a, const NbKey& b) {
    return a.rank == b.rank && a.x == b.x && a.y == b.y;
}

// Block-wide exclusive prefix sum of one int per thread; returns the exclusive prefix and the
// block total.  `warp_sums` is >= 32 ints of shared memory.  All threads of the block must call.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // protect warp_sums reuse across consecutive calls
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarps ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;            // exclusive warp offsets
        if (lane == 31) warp_sums[32] = winc;  // block total
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[warp] + inc - v;
}

}  // namespace psim
