// psim_tiled.cu -- the "tiled" engine: persistent tile-resident particles, one fused kernel per step.
//
// Layout in HBM.  The box is cut into square tiles of TS x TS cutoff cells.  Every tile owns a stripe
// of CAP particle slots in three vectorised streams -- pos (double2 x,y), vel (double2 vx,vy), id (int)
// -- plus acc (double2 ax,ay), written only on steps whose accelerations are kept.  A tile's live
// particles occupy slots [0, count).  The stripes and the tile "exports" (eight halo lists with the
// positions of the particles in the tile's boundary cells, and an outbox with the full records of the
// particles that left the tile) are double buffered by step parity: a step reads parity p and writes
// p^1, so no CTA ever reads what another CTA of the same launch writes.  Steady-state HBM traffic is
// read 36 B + write 36 B per particle-step plus a few percent of exports: there is no global histogram,
// scan or scatter in the time loop.
//
// Per step one persistent kernel (tile_step_kernel); a CTA walks tiles blockIdx.x, +gridDim.x, ...
// Seven consumer warps plus one PRODUCER warp that sleeps in the hardware barriers the consumers pass anyway:
//   producer, one tile ahead: list counts fetched early, then TMA bulk copies (cp.async.bulk + mbarrier) into the
//      other shared-memory stage -- the tile's three stripes, the eight neighbour halo lists (landing directly
//      behind the own particles as the one-cell "apron") and the nine surrounding outboxes packed back to back --
//      then the INGEST: particles that entered the tile last step (outbox records whose destination cell lies
//      in the tile) are appended to the stripe, records in the apron ring become apron particles, and apron +
//      migrants + the own slots past the consumers' share are binned.
//   A. bin (consumers, during the PREVIOUS tile's phase C): every particle that was already in the tile goes
//      into a (TS+2)^2 cell table IN SHARED MEMORY -- per cell the head of a linked list (one atomicExch per
//      particle; reference part3/gpu.cu:92-112 does this in global memory with 16 fixed slots per cell) and one
//      bit in a per-row occupancy map; its position in cell units is kept in FP32.      -- barrier 1 --
//   B. search: every own particle takes the 9 occupancy bits of its 3x3 neighbourhood (reference
//      part1/serial.cpp:102-117), walks only the non-empty cells and tests candidates in FP32 against a slightly
//      widened cutoff; survivors go to a CTA-wide pair list.                                -- barrier 2 --
//   C. pair evaluation, dense (last consumer warp): one lane per listed pair redoes the distance test in exact
//      FP64 and, if in range, evaluates the coefficient (sqrt + three divisions, reference part1/serial.cpp:19-36):
//      the expensive sequence runs once per tile with full lanes instead of once per warp with ~3.  Particles with
//      three or more candidates take the exact canonical-order path here.  The other warps run phase A of the next
//      tile meanwhile.                                                                      -- barrier 3 --
//   D+E. sum (<= 2 contributions are order independent), move + reflect (reference part1/serial.cpp:46-61), new
//      cell, stay / leave; slots come from one warp-aggregated shared atomic per warp, so without a further
//      barrier stayers are stored compacted into the other stripe buffer, leavers into the tile's outbox and
//      particles in boundary cells into the edge / corner halo lists the neighbours read next step.
//
// Slabs (precedent reference part2/mpi.cpp:258-270,296-365): a rank owns a contiguous range
// of tile rows plus one ghost tile row on each side that holds only exports.  The exports of one tile row are one
// contiguous byte range, so the halo exchange AND the particle migration between GPUs are the same data: either one
// ncclSend/ncclRecv of the first / last owned row per neighbour, or -- the default on one NVSwitch box -- stored by
// the boundary-row launch straight into the neighbour's ghost row over NVLink (tiled_step).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "psim_force.cuh"
#include "psim_internal.h"
#include "psim_tiled.h"

namespace psim {

// ------------------------------------------------------------------------------------------
// compile-time tile configurations
// ------------------------------------------------------------------------------------------
// CAP slots per tile (mean population is 0.2*TS^2), HE / HC entries per edge / corner halo list,
// CO outbox records per tile; the outboxes of the 3x3 tiles are staged in shared memory back to back, OS
// records in total (a tile whose surroundings hold more reads the rest straight from global memory: this
// only happens in bursts, e.g. a column of the initial lattice that sits exactly on a tile boundary),
// NP pair-list entries, THREADS consumer threads per CTA (every CTA has one more warp, the producer),
// CTAS resident per SM.  THREADS is a little above the MEAN population, so nearly every lane has a
// particle in the first pass; the second pass (PER = 2) only runs for the slots past THREADS.
template <int TS> struct TileCfg;
template <> struct TileCfg<16> { static constexpr int CAP = 128,  HE = 24, HC = 8, CO = 64, OS = 24, NP = 32,  THREADS = 64,  CTAS = 8; };
template <> struct TileCfg<32> { static constexpr int CAP = 352,  HE = 32, HC = 8, CO = 64, OS = 48, NP = 64,  THREADS = 224, CTAS = 4; };
template <> struct TileCfg<64> { static constexpr int CAP = 1152, HE = 64, HC = 8, CO = 64, OS = 96, NP = 320, THREADS = 832, CTAS = 1; };

struct __align__(16) OutRec {  // one migrating particle, 64 bytes
    double x, y, vx, vy, ax, ay;
    int id;
    int row, col;  // its cell (row = floor(x/0.01), col = floor(y/0.01)), computed once by the sender
    int pad;
};
static_assert(sizeof(OutRec) == 64, "OutRec must be 64 bytes");

constexpr unsigned kEmpty = 0xFFFFu;          // end of a cell's list / empty cell
constexpr float kPrefilter2 = 1.001f;         // FP32 candidate test in cell units: cutoff^2 = 1 widened by 1e-3 (float error < 2e-5)

template <int TS> struct TileDims {
    using C = TileCfg<TS>;
    static constexpr int W = TS + 2, NC = W * W;
    static constexpr int RW = (W + 31) / 32 + 1;          // occupancy words per cell row (+1: funnel shifts read one past)
    static constexpr int NC4 = (NC + 3) / 4 * 4, RB4 = (W * RW + 3) / 4 * 4;   // table sizes padded to 16 bytes
    static constexpr int HL = 4 * C::HE + 4 * C::HC;      // halo entries a tile exports / stages
    static constexpr int MAXH = HL + 16;                  // apron capacity (halo lists + apron outbox records)
    static constexpr int PTOT = C::CAP + MAXH;
    static constexpr int PER = (C::CAP + C::THREADS - 1) / C::THREADS;
    static constexpr int NW = C::THREADS / 32;            // consumer warps
    static constexpr int OS = C::OS;                      // staged outbox records (all nine outboxes together)
    static_assert(C::CAP % 4 == 0 && C::THREADS % 32 == 0 && NW <= 32 && PTOT < 0xFFFF, "configuration");
};

// one pipeline stage in shared memory (TMA destinations, 16-byte aligned)
template <int TS> struct __align__(16) Stage {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    double2 xy[D::PTOT];    // own [0, n) then apron [CAP, CAP + n_apron)
    double2 v[C::CAP];
    int id[C::CAP];
    OutRec obox[D::OS];     // staged records of the outboxes of the 3x3 tiles, contiguous
};

template <int TS> struct __align__(16) TileSmem {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    Stage<TS> st[2];
    double2 pres[C::NP];                      // pair list: contribution of pair t to its first particle
    float2 rel[D::PTOT];                      // tile-relative FP32 positions in cell units (prefilter only; read before barrier 2,
                                              // the producer writes the next tile's after it)
    // per cell: first particle of its list (kEmpty: none); per cell row: occupancy bit per cell.  Double buffered by tile,
    // padded to whole 16-byte words so that a table is wiped with 128-bit stores
    alignas(16) unsigned head[2][D::NC4];
    alignas(16) unsigned rowbits[2][D::RB4];
    unsigned pij[C::NP];                      // pair list: i | j << 16
    unsigned long long full[2];               // mbarriers: stage filled by TMA
    unsigned short next[2][D::PTOT];          // next particle in the same cell (double buffered like head: the exact path of
                                              // a slow warp may still walk the lists while a fast warp bins the next tile)
    unsigned short pcell[D::PTOT];            // local row << 8 | local column of a binned particle (read before barrier 2, the producer writes the next
                                              // tile's after it)
    unsigned short pcode[C::CAP];             // per own particle: search result (pair count | pair-list base << 2), B -> D
    int cnts[2][32];                          // staged tiles: [0] own, [1..8] halo lists, [9..17] outboxes, [18] halo total,
                                              // [19] staged outbox records, [20..28] staged records per outbox
    int hout[8];
    int n_own[2];                             // per stage: population after ingest
    int prev_lr, prev_tc;                     // tile whose export counts still have to be published
    int npairs, flags;
    int n_stay, n_leave;                      // stayers / leavers of the current tile (warp-aggregated atomics)
    int hw_leave, hw_halo, hw_tile, hw_apron, hw_pairs;   // high-water marks over the tiles this CTA processed
};

static_assert(4 * (sizeof(TileSmem<32>) + 1024) <= 233472, "four CTAs of the 32-cell tile kernel must fit one SM's shared memory");

__host__ __device__ inline int halo_offset(int list, int HE, int HC) { return list < 4 ? list * HE : 4 * HE + (list - 4) * HC; }
__host__ __device__ inline int halo_cap(int list, int HE, int HC) { return list < 4 ? HE : HC; }

struct TileParams {
    const double2 *pos_in, *vel_in;
    const int* id_in;
    double2 *pos_out, *vel_out, *acc_out;
    int* id_out;
    int* tcount;
    const char* exp_in;
    char* exp_out;
    ExportLayout L;
    int ntx, nty;    // tiles per side (global)
    int tr_base;     // global tile row of local row 0
    int lrow0;       // first local tile row of this launch
    int row_stride;  // tile t of the launch lies in local row lrow0 + (t / ntx) * row_stride (boundary launch: first and last row)
    int ntiles;      // tiles of this launch (rows * ntx)
    int bincnt;
    double size;
    int* err;
    // slabs with peer-memory exchange: the exports of the first / last owned tile row are ALSO written straight into
    // the neighbour GPU's ghost row (NVLink peer stores), so the step needs no separate exchange
    char* peer_row[2];   // [0] lower neighbour's upper ghost row, [1] upper neighbour's lower ghost row (parity written); or null
    int last_lrow;       // local index of the last owned tile row
};

__device__ __forceinline__ const char* row_ptr(const char* base, const ExportLayout& L, int lrow) {
    return base + (size_t)lrow * L.row_bytes;
}
__device__ __forceinline__ char* row_ptr(char* base, const ExportLayout& L, int lrow) {
    return base + (size_t)lrow * L.row_bytes;
}
// the neighbour's ghost row that mirrors local row lr (null for interior rows and without peer exchange)
__device__ __forceinline__ char* peer_row_of(const TileParams& P, int lr) {
    return lr == 1 ? P.peer_row[0] : (lr == P.last_lrow ? P.peer_row[1] : nullptr);
}

// Apron source k = 0..7: which neighbour (dr, dc) and which of ITS lists faces this tile.
//   k: 0 N-neighbour's S list | 1 S-neighbour's N | 2 W-neighbour's E | 3 E-neighbour's W
//      4 NW-neighbour's SE corner | 5 NE's SW | 6 SW's NE | 7 SE's NW
// list ids: 0 N, 1 S, 2 W, 3 E, 4 NW, 5 NE, 6 SW, 7 SE (N = row 0 of the tile, W = column 0)
__device__ __forceinline__ void halo_source(int k, int& dr, int& dc, int& list) {
    // packed tables, 4 bits per entry: dr+1, dc+1, list
    dr = (int)((0x22001120u >> (4 * k)) & 0xF) - 1;    // k0..7: -1 1 0 0 -1 -1 1 1  -> +1: 0 2 1 1 0 0 2 2
    dc = (int)((0x20202011u >> (4 * k)) & 0xF) - 1;    // k0..7:  0 0 -1 1 -1 1 -1 1 -> +1: 1 1 0 2 0 2 0 2
    list = (int)((0x45672301u >> (4 * k)) & 0xF);      // k0..7:  1 0 3 2 7 6 5 4
}

// ---- PTX helpers: mbarrier + 1-D bulk copy (TMA) --------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// consumers: the data is normally there already, spin tightly
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// producer: waits for a whole tile of consumer work; let the hardware suspend the warp (time hint in ns)
// so that the wait does not steal issue slots from the consumers
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier `ID` among N threads (id 0 is __syncthreads)
template <int ID, int N>
__device__ __forceinline__ void named_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tile index of this launch -> local row / column
struct TileCoord {
    int lr, tc, tr, lt;
};
__device__ __forceinline__ TileCoord tile_coord(const TileParams& P, int t) {
    TileCoord c;
    c.lr = P.lrow0 + (t / P.ntx) * P.row_stride;
    c.tc = t % P.ntx;
    c.tr = P.tr_base + c.lr;
    c.lt = c.lr * P.ntx + c.tc;
    return c;
}

// walks the tiles first, first + G, first + 2G, ... of a launch without a division per tile
struct TileWalker {
    int lr, tc, gq, gr;
    __device__ __forceinline__ void init(const TileParams& P, int first, int G) {
        lr = P.lrow0 + (first / P.ntx) * P.row_stride;
        tc = first % P.ntx;
        gq = (G / P.ntx) * P.row_stride;
        gr = G % P.ntx;
    }
    __device__ __forceinline__ void advance(const TileParams& P) {
        tc += gr;
        lr += gq;
        if (tc >= P.ntx) {
            tc -= P.ntx;
            lr += P.row_stride;
        }
    }
    __device__ __forceinline__ TileCoord coord(const TileParams& P) const {
        TileCoord c;
        c.lr = lr;
        c.tc = tc;
        c.tr = P.tr_base + lr;
        c.lt = lr * P.ntx + tc;
        return c;
    }
};

// list count j of tile t: j = 0 own population, 1..8 apron sources, 9..17 outboxes of the 3x3 tiles.
// Returns the RAW stored value (clamp_count applies the capacity later) so that nothing depends on the
// global load until the caller needs the number.
template <int TS>
__device__ __forceinline__ int clamp_count(int raw, int j) {
    using C = TileCfg<TS>;
    return min(raw, j == 0 ? C::CAP : j >= 9 ? C::CO : j <= 4 ? C::HE : C::HC);
}
template <int TS>
__device__ __forceinline__ int load_count(const TileParams& P, const TileCoord& c, int j) {
    if (j == 0) return P.tcount[c.lt];
    int dr, dc, list;
    if (j <= 8) {
        halo_source(j - 1, dr, dc, list);
    } else {
        dr = (j - 9) / 3 - 1;
        dc = (j - 9) % 3 - 1;
        list = 8;
    }
    const int ntr = c.tr + dr, ntc = c.tc + dc;
    if (ntr < 0 || ntr >= P.nty || ntc < 0 || ntc >= P.ntx) return 0;
    const int* ec = reinterpret_cast<const int*>(row_ptr(P.exp_in, P.L, c.lr + dr) + P.L.off_cnt) + (size_t)ntc * 16;
    return ec[list];
}

// Loader (one warp): given the list counts of a tile (lane j holds count j, fetched earlier so that the
// global-load latency is hidden), publish them in `cnt` and issue the bulk copies into `st`.
//   copy 0 pos, 1 vel, 2 id, 3..10 apron source k = copy-3 (straight to its final place behind the own
//   particles), 11..19 outbox nb = copy-11 (first CS records, packed back to back)
struct TileCopy {   // what one producer lane copies for a tile: prepared one tile early, fired the moment the stage is free
    const void* src;
    void* dst;
    unsigned bytes, total;
};
template <int TS>
__device__ __forceinline__ TileCopy prepare_tile(const TileParams& P, const TileCoord c, Stage<TS>& st, int* cnt, int cj_raw, int lane) {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    const int cj = lane < 18 ? clamp_count<TS>(cj_raw, lane) : 0;
    // exclusive prefixes in ONE scan: halo entries over lanes 1..8 (low 16 bits), outbox records over lanes 9..17 (high)
    const int hv = (lane >= 1 && lane <= 8) ? cj : 0;
    const int ov = (lane >= 9 && lane <= 17) ? cj : 0;
    int scan = hv | (ov << 16);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, scan, o);
        if (lane >= o) scan += a;
    }
    const int hs = scan & 0xFFFF, os = scan >> 16;
    const int totals = __shfl_sync(0xffffffffu, scan, 17);
    const int n_halo = totals & 0xFFFF, n_staged = min(totals >> 16, D::OS);
    const int staged = max(0, min(ov, D::OS - (os - ov)));   // lanes 9..17: records of my outbox that fit the staging area
    const int n = __shfl_sync(0xffffffffu, cj, 0);
    // what this lane copies: count, destination offset and staged records of "its" list, fetched from the lane that holds them
    const int src_lane = lane < 3 ? 0 : lane < 20 ? lane - 2 : 0;
    const int mine = __shfl_sync(0xffffffffu, cj | (staged << 16), src_lane);
    const int offs = __shfl_sync(0xffffffffu, (hs - hv) | ((os - ov) << 16), src_lane);
    const int my_cnt = mine & 0xFFFF, my_staged = mine >> 16, my_hoff = offs & 0xFFFF, my_ooff = offs >> 16;
    if (lane < 18) cnt[lane] = cj;
    if (lane == 18) cnt[18] = n_halo;
    if (lane == 19) cnt[19] = n_staged;
    if (lane >= 9 && lane <= 17) cnt[11 + lane] = staged;
    const size_t gbase = (size_t)c.lt * C::CAP;
    const void* src = nullptr;
    void* dst = nullptr;
    unsigned bytes = 0;
    if (lane == 0) {
        bytes = (unsigned)n * 16u; src = P.pos_in + gbase; dst = st.xy;
    } else if (lane == 1) {
        bytes = (unsigned)n * 16u; src = P.vel_in + gbase; dst = st.v;
    } else if (lane == 2) {
        bytes = (unsigned)((n + 3) >> 2) * 16u; src = P.id_in + gbase; dst = st.id;
    } else if (lane < 11) {
        int dr, dc, list;
        halo_source(lane - 3, dr, dc, list);
        bytes = (unsigned)my_cnt * 16u;
        if (bytes) {
            src = reinterpret_cast<const double2*>(row_ptr(P.exp_in, P.L, c.lr + dr) + P.L.off_hxy) +
                  (size_t)(c.tc + dc) * D::HL + halo_offset(list, C::HE, C::HC);
            dst = st.xy + C::CAP + my_hoff;
        }
    } else if (lane < 20) {
        const int nb = lane - 11;
        const int dr = nb / 3 - 1, dc = nb % 3 - 1;
        bytes = (unsigned)my_staged * (unsigned)sizeof(OutRec);
        if (bytes) {
            src = reinterpret_cast<const OutRec*>(row_ptr(P.exp_in, P.L, c.lr + dr) + P.L.off_obox) + (size_t)(c.tc + dc) * C::CO;
            dst = st.obox + my_ooff;
        }
    }
    // bytes in flight for this tile, known without a reduction: two 16-byte stripes, the id stripe rounded up to 16 bytes, the
    // halo entries and the staged outbox records
    const unsigned total = (unsigned)n * 32u + (unsigned)((n + 3) >> 2) * 16u + (unsigned)n_halo * 16u +
                           (unsigned)n_staged * (unsigned)sizeof(OutRec);
    TileCopy k;
    k.src = src;
    k.dst = dst;
    k.bytes = bytes;
    k.total = total;
    return k;
}
__device__ __forceinline__ void fire_tile(const TileCopy& k, unsigned long long* bar, int lane) {
    __syncwarp();   // counts are in shared memory before the barrier can complete
    if (lane == 0) mbar_arrive_expect_tx(bar, k.total);
    __syncwarp();
    if (k.bytes) tma_load_1d(k.dst, k.src, k.bytes, bar);
}

// Rare path: exact FP64 walk of the 3x3 neighbourhood with canonical-order summation (particles whose
// prefilter found three or more candidates, or whose pairs did not fit the pair list).
template <int W>
static __device__ __noinline__ double2 slow_force(const double2* xy, const unsigned* head, const unsigned short* next, int i,
                                                  int cell) {
    const double2 pi = xy[i];
    auto visit = [&](auto&& f) {
#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
            unsigned h = head[cell + (k / 3 - 1) * W + (k % 3 - 1)];
            while (h != kEmpty) {
                const double2 pj = xy[h];
                f(pj.x, pj.y, k);
                h = next[h];
            }
        }
    };
    auto rank_of = [&](double, double, int k) { return visit_rank(k / 3 - 1, k % 3 - 1); };
    double ax, ay;
    int nb;
    accumulate_force(pi.x, pi.y, visit, rank_of, ax, ay, nb);
    return make_double2(ax, ay);
}

#ifdef PSIM_DEBUG_GUARDS
// debug build only: bounds guards on every data-dependent shared-memory index; a hit sets a bit in err[5]
#define PSIM_GUARD(cond, bit, ...)                 \
    if (!(cond)) {                                 \
        atomicOr(P.err + 5, 1 << (bit));           \
        atomicMax(P.err + 6, first + it * G);      \
        __VA_ARGS__;                               \
    }
#else
#define PSIM_GUARD(cond, bit, ...)
#endif

#ifdef PSIM_PHASE_TIMERS
// profiling build only: cycles spent per phase, summed over tiles, for warp 0, a middle warp and the loader warp
__device__ unsigned long long g_phase_cycles[34][12];
__device__ unsigned g_hist[8][64];   // histograms (256-cycle bins) of per-warp per-tile durations: 0 A, 1 B, 2 DE, 3 top+wait
#define PSIM_HIST(h, k) do { if (lane == 0) atomicAdd(&g_hist[h][min((tacc[k] - hprev[k]) >> 8, 63u)], 1u); hprev[k] = tacc[k]; } while (0)
#define PSIM_HIST_KEEP(h, k) do { if (lane == 0) atomicAdd(&g_hist[h][min((tacc[k] - hprev[k]) >> 8, 63u)], 1u); } while (0)
#define PSIM_TICK(k)                                   \
    do {                                               \
        /* a barrier only blocks at the next access to shared memory: force one before reading the clock */ \
        asm volatile("" ::"r"(*(volatile int*)&S.flags) : "memory"); \
        const unsigned now__ = (unsigned)clock();      \
        tacc[k] += now__ - tlast;                      \
        tlast = now__;                                 \
    } while (0)
#else
#define PSIM_TICK(k) do {} while (0)
#define PSIM_HIST(h, k) do {} while (0)
#define PSIM_HIST_KEEP(h, k) do {} while (0)
#endif

// ------------------------------------------------------------------------------------------
// the per-step kernel
// ------------------------------------------------------------------------------------------
// Bin particle p of stage sb: exact cell, FP32 position in cell units relative to the table origin (cell
// (r0-1, c0-1); rounding error < 4e-6 cells),
// list head exchange, occupancy bit.  `ring_check`: apron entries may lie outside the table (never for the
// lists of proper neighbours; kept as a memory-safety guard).
template <int TS, bool kRingCheck>
__device__ __forceinline__ void bin_particle(TileSmem<TS>& S, int sb, int p, double x, double y, int r0m1, int c0m1, int bincnt) {
    using D = TileDims<TS>;
    const double qx = div_by_bin(x), qy = div_by_bin(y);   // exact position in cell units
    const int lrow = min(max(__double2int_rd(qx), 0), bincnt - 1) - r0m1, lcol = min(max(__double2int_rd(qy), 0), bincnt - 1) - c0m1;
#ifdef PSIM_DEBUG_GUARDS
    if ((unsigned)lrow >= (unsigned)D::W || (unsigned)lcol >= (unsigned)D::W || (unsigned)p >= (unsigned)D::PTOT) { atomicOr(&S.flags, 1 << 20); return; }
#endif
    if (kRingCheck && ((unsigned)lrow >= (unsigned)D::W || (unsigned)lcol >= (unsigned)D::W)) return;
    const int cell = lrow * D::W + lcol;
    const unsigned old = atomicExch(&S.head[sb][cell], (unsigned)p);
    S.next[sb][p] = (unsigned short)old;
    S.pcell[p] = (unsigned short)((lrow << 8) | lcol);   // row and column separately: no division when it is read back
    atomicOr(&S.rowbits[sb][lrow * D::RW + (lcol >> 5)], 1u << (lcol & 31));
    S.rel[p] = make_float2(__double2float_rn(__dsub_rn(qx, (double)r0m1)), __double2float_rn(__dsub_rn(qy, (double)c0m1)));
}

template <int TS, bool kStoreAcc, bool kPeer>
__global__ void __launch_bounds__(TileCfg<TS>::THREADS + 32, TileCfg<TS>::CTAS) tile_step_kernel(const TileParams P) {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    constexpr int T = C::THREADS, CAP = C::CAP, W = D::W, RW = D::RW, NW = D::NW;
    constexpr int HE = C::HE, HC = C::HC, CO = C::CO, NP = C::NP;
    // named barriers: 1 = cell table complete (consumers + producer), 2 = pair list complete (consumers),
    // 3 = pair contributions ready and the other cell table clean (consumers + producer)
    constexpr int kBarTable = 1, kBarPairs = 2, kBarForces = 3;
    constexpr int TA = (NW - 1) * 32;   // consumer threads that bin (all warps but the last, which evaluates the pairs)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem<TS>& S = *reinterpret_cast<TileSmem<TS>*>(smem_raw);

    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));   // volatile: keep it in a register instead of re-reading the special register
    const int lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int first = blockIdx.x;
    if (first >= P.ntiles) return;

    // ---- prologue ---------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.npairs = 0;
        S.flags = 0;
        S.n_stay = S.n_leave = 0;
        S.hw_leave = S.hw_halo = S.hw_tile = S.hw_apron = S.hw_pairs = 0;
        S.prev_lr = -1;
        S.prev_tc = 0;
    }
    if (tid < 8) S.hout[tid] = 0;
    for (int c = tid; c < 2 * D::NC4; c += T + 32) (&S.head[0][0])[c] = kEmpty;
    for (int c = tid; c < 2 * D::RB4; c += T + 32) (&S.rowbits[0][0])[c] = 0u;
    __syncthreads();

    TileWalker walk;
    walk.init(P, first, G);

    // =================================================================================================
    // producer warp: feeds the two-stage pipeline one tile ahead, ingests the migrants and bins the apron of
    // the next tile.  It sleeps in the hardware barriers the consumers pass anyway, so it costs no issue
    // slots while idle.
    // =================================================================================================
    if (warp == NW) {
        const unsigned lt_mask = (1u << lane) - 1u;
        // Ingest: outbox records of the 3x3 tiles around tile c (staged in stage sb).  Records now inside the
        // tile are appended to its stripe, records in its apron ring become apron particles (after the
        // halo-list entries).  One lane per record.  Runs while the consumers still work on the previous tile.
        auto ingest = [&](int sb, const TileCoord c, unsigned full_parity) {
            Stage<TS>& st = S.st[sb];
            const int* cnt = S.cnts[sb];
            mbar_wait_relaxed(&S.full[sb], full_parity);
            const int r0 = c.tr * TS, c0 = c.tc * TS;
            int n_own = cnt[0], n_ap = cnt[18], flags = 0;
            auto take = [&](bool valid, const OutRec* rec) {
                bool mine = false, apron = false;
                if (valid) {
                    const int lrow = rec->row - (r0 - 1), lcol = rec->col - (c0 - 1);
                    const bool ring = lrow >= 0 && lrow <= TS + 1 && lcol >= 0 && lcol <= TS + 1;
                    mine = lrow >= 1 && lrow <= TS && lcol >= 1 && lcol <= TS;
                    apron = ring && !mine;
                }
                const unsigned mm = __ballot_sync(0xffffffffu, mine), am = __ballot_sync(0xffffffffu, apron);
                if (mine | apron) {
                    const int d = mine ? n_own + __popc(mm & lt_mask) : CAP + n_ap + __popc(am & lt_mask);
                    if (d < (mine ? CAP : D::PTOT)) {
                        const double x = rec->x, y = rec->y;
                        st.xy[d] = make_double2(x, y);
                        if (mine) {
                            st.v[d] = make_double2(rec->vx, rec->vy);
                            st.id[d] = rec->id;
                        }
                        bin_particle<TS, true>(S, sb, d, x, y, r0 - 1, c0 - 1, P.bincnt);
                    }
                }
                n_own += __popc(mm);
                n_ap += __popc(am);
            };
            const int n_staged = cnt[19];
#pragma unroll 1
            for (int q0 = 0; q0 < n_staged; q0 += 32) take(q0 + lane < n_staged, &st.obox[min(q0 + lane, D::OS - 1)]);
            if (n_staged == D::OS) {
                // bursts only: records that did not fit the staging area are read from the neighbours' outboxes in global memory
#pragma unroll 1
                for (int nb = 0; nb < 9; ++nb) {
                    const int n = cnt[9 + nb], s0 = cnt[20 + nb];
                    if (n <= s0) continue;
                    const OutRec* grec = reinterpret_cast<const OutRec*>(row_ptr(P.exp_in, P.L, c.lr + nb / 3 - 1) + P.L.off_obox) +
                                         (size_t)(c.tc + nb % 3 - 1) * CO;
                    for (int e0 = s0; e0 < n; e0 += 32) take(e0 + lane < n, grec + min(e0 + lane, n - 1));
                }
            }
            // own particles past the slots the consumer warps bin themselves (one slot per binning thread)
            {
                const int n_own0 = cnt[0];
#pragma unroll 1
                for (int p = TA + lane; p < n_own0; p += 32) {
                    const double2 a = st.xy[p];
                    bin_particle<TS, false>(S, sb, p, a.x, a.y, r0 - 1, c0 - 1, P.bincnt);
                }
            }
            // the halo-list apron (already in place behind the own particles)
            {
                const int n_halo = cnt[18];
#pragma unroll 1
                for (int q = lane; q < n_halo; q += 32) {
                    const double2 a = st.xy[CAP + q];
                    bin_particle<TS, true>(S, sb, CAP + q, a.x, a.y, r0 - 1, c0 - 1, P.bincnt);
                }
            }
            if (n_own > CAP) { flags |= kErrTileOverflow; n_own = CAP; }
            if (n_ap > D::MAXH) { flags |= kErrSmemOverflow; n_ap = D::MAXH; }
            if (lane == 0) {
                S.n_own[sb] = n_own;
                S.hw_tile = max(S.hw_tile, n_own);
                S.hw_apron = max(S.hw_apron, n_ap);
                if (flags) atomicOr(&S.flags, flags);
            }
            fence_proxy_async();   // my generic writes to this stage are ordered before the bulk copies that refill it
            __syncwarp();
        };

        // The copies of a tile are PREPARED (list counts, offsets, addresses, byte total) one tile before they are FIRED, so
        // that the bulk copies start the moment barrier 1 frees the stage: their latency, not the producer's arithmetic,
        // then decides when the consumers can bin the next tile.
        const TileCoord c0 = walk.coord(P);
        fire_tile(prepare_tile<TS>(P, c0, S.st[0], S.cnts[0], lane < 18 ? load_count<TS>(P, c0, lane) : 0, lane), &S.full[0], lane);
        walk.advance(P);   // -> tile 1
        TileCopy next_copy{nullptr, nullptr, 0u, 0u};
        if (first + G < P.ntiles)
            next_copy = prepare_tile<TS>(P, walk.coord(P), S.st[1], S.cnts[1], lane < 18 ? load_count<TS>(P, walk.coord(P), lane) : 0, lane);
        ingest(0, c0, 0u);
#ifdef PSIM_PHASE_TIMERS
        unsigned tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        unsigned tlast = (unsigned)clock();
#endif
        for (int it = 0;; ++it) {
            const int t = first + it * G;
            if (t >= P.ntiles) break;
            const int sb = it & 1;
            const bool has_next = t + G < P.ntiles, has_next2 = t + 2 * G < P.ntiles;
            const TileCoord cn = walk.coord(P);   // tile it + 1
            walk.advance(P);                      // -> tile it + 2
            const TileCoord cn2 = walk.coord(P);
            int cj2 = 0;
            if (has_next2 && lane < 18) cj2 = load_count<TS>(P, cn2, lane);   // in flight across the barrier
            PSIM_TICK(0);
            named_sync<kBarTable, T + 32>();   // every consumer warp has left tile it - 1: its stage is free
            PSIM_TICK(1);
            if (has_next) fire_tile(next_copy, &S.full[sb ^ 1], lane);
            // tile it + 2 will use THIS tile's stage: only its addresses and count words are set up now (nobody reads the count
            // words of a tile once it is binned and ingested), the copies are fired after the next barrier 1
            if (has_next2) next_copy = prepare_tile<TS>(P, cn2, S.st[sb], S.cnts[sb], cj2, lane);
            PSIM_TICK(2);
            named_sync<kBarForces, T + 32>();  // the consumers have cleaned the other cell table and are done with rel / pcell
            PSIM_TICK(3);
            if (has_next) ingest(sb ^ 1, cn, (unsigned)(((it + 1) >> 1) & 1));
            PSIM_TICK(4);
        }
#ifdef PSIM_PHASE_TIMERS
        if (lane == 0) for (int k = 0; k < 9; ++k) atomicAdd(&g_phase_cycles[33][k], (unsigned long long)tacc[k]);
#endif
        return;
    }

    // =================================================================================================
    // consumer warps
    // =================================================================================================
    const unsigned lt_mask = (1u << lane) - 1u;
    // Deferred bookkeeping of the tile that was just finished, run by nine threads right after a barrier
    // (all appends / slot reservations of that tile happened before it): tid < 8 publishes halo list `tid`,
    // tid == 8 the outbox count and the tile's new population.
    auto publish_counts = [&]() {
        const int done_lr = S.prev_lr, done_tc = S.prev_tc;
        if (done_lr < 0) return;
        int* ec = reinterpret_cast<int*>(row_ptr(P.exp_out, P.L, done_lr) + P.L.off_cnt) + (size_t)done_tc * 16;
        char* prow = kPeer ? peer_row_of(P, done_lr) : nullptr;
        int* pec = prow ? reinterpret_cast<int*>(prow + P.L.off_cnt) + (size_t)done_tc * 16 : nullptr;
        if (tid < 8) {
            const int k = S.hout[tid], cap = halo_cap(tid, HE, HC);
            if (k > cap) atomicOr(&S.flags, kErrHaloOverflow);
            ec[tid] = min(k, cap);
            if (pec) pec[tid] = min(k, cap);
            S.hout[tid] = 0;
            if (tid < 4) atomicMax(&S.hw_halo, k);
        } else {
            const int n_stay = S.n_stay, n_leave = S.n_leave;
            S.n_stay = 0;
            S.n_leave = 0;
            if (n_leave > CO) atomicOr(&S.flags, kErrOutboxOverflow);
            ec[8] = min(n_leave, CO);
            if (pec) pec[8] = min(n_leave, CO);
            P.tcount[done_lr * P.ntx + done_tc] = n_stay;
            S.hw_leave = max(S.hw_leave, n_leave);
        }
    };

#ifdef PSIM_PHASE_TIMERS
    unsigned tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned hprev[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned tlast = (unsigned)clock();
#endif
    // Phase A of a tile -- binning the particles that were already there, one per thread of the first NW-1 warps (the
    // producer bins the rest, the migrants and the apron) -- runs during the PREVIOUS tile's pair-evaluation phase,
    // when these warps would otherwise idle; only the first tile is binned up front.
    auto bin_own = [&](int sb_, unsigned full_parity, int r0_, int c0_) {
        if (warp < NW - 1) {
            mbar_wait(&S.full[sb_], full_parity);   // the tile's bytes (and counts) have landed
            if (tid < S.cnts[sb_][0]) {
                const double2 a = S.st[sb_].xy[tid];
                bin_particle<TS, false>(S, sb_, tid, a.x, a.y, r0_ - 1, c0_ - 1, P.bincnt);
            }
        }
    };
    bin_own(0, 0u, (P.tr_base + walk.lr) * TS, walk.tc * TS);
    for (int it = 0;; ++it) {
        if (first + it * G >= P.ntiles) break;
        const int sb = it & 1;   // stage and cell-table buffer of this tile
        Stage<TS>& st = S.st[sb];
        const unsigned* head = S.head[sb];
        const unsigned short* next = S.next[sb];
        const int lr = walk.lr, tc = walk.tc, tr = P.tr_base + lr;
        const int r0 = tr * TS, c0 = tc * TS;

        PSIM_TICK(0);
        PSIM_TICK(1);
        PSIM_TICK(2);
        named_sync<kBarTable, T + 32>();   // (1) cell table complete; every warp has left the previous tile
        PSIM_TICK(3);
        if (tid < 9) publish_counts();
        const int n_own = S.n_own[sb];
        // the other cell table (last read before this barrier) is cleaned now: the next tile is binned into it during
        // this tile's pair-evaluation phase
        {
            uint4* oh = reinterpret_cast<uint4*>(S.head[sb ^ 1]);
            uint4* ob = reinterpret_cast<uint4*>(S.rowbits[sb ^ 1]);
#pragma unroll 1
            for (int c = tid; c < D::NC4 / 4; c += T) oh[c] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            if (tid < D::RB4 / 4) ob[tid] = make_uint4(0u, 0u, 0u, 0u);
            static_assert(D::RB4 / 4 <= T, "one 128-bit store per thread wipes the occupancy map");
        }

        // ---- B: candidate search in FP32 ------------------------------------------------------------------
#pragma unroll 1
        for (int i = tid; i < n_own; i += T) {
            const int rc = S.pcell[i];
            const int lrow = rc >> 8, lcol = rc & 0xFF;
            const int cell = lrow * W + lcol;
            PSIM_GUARD(lrow >= 1 && lrow <= TS && lcol >= 1 && lcol <= TS, 0, continue)
            const float2 ri = S.rel[i];
            // 9 occupancy bits of the 3x3 neighbourhood, bit 3*(dr+1) + (dc+1)
            const unsigned* rb = S.rowbits[sb] + (lrow - 1) * RW + ((lcol - 1) >> 5);
            const unsigned sh = (unsigned)(lcol - 1) & 31u;
            unsigned m = 0;
#pragma unroll
            for (int dr = 0; dr < 3; ++dr) m |= (__funnelshift_r(rb[dr * RW], rb[dr * RW + 1], sh) & 7u) << (3 * dr);
            if (head[cell] == (unsigned)i && next[i] == kEmpty) m &= ~16u;   // alone in my own cell
            int fc = 0;
            unsigned cand = 0;   // the last two prefilter hits, 16 bits each
            while (m) {
                const int k = __ffs((int)m) - 1;
                m &= m - 1;
                const int q = (k * 11) >> 5;   // k / 3
                unsigned h = head[cell + q * (W - 3) + k - (W + 1)];
                do {   // the occupancy bit guarantees a non-empty list
                    PSIM_GUARD(h < (unsigned)D::PTOT, 1, break)
                    const float2 rj = S.rel[h];
                    const unsigned hn = next[h];
                    const float dx = rj.x - ri.x, dy = rj.y - ri.y;
                    const float r2 = dx * dx + dy * dy;
                    if (r2 <= kPrefilter2 && h != (unsigned)i) {
                        cand = (cand << 16) | h;
                        ++fc;
                    }
                    h = hn;
                } while (h != kEmpty);
            }
            // code: fc (0, 1, 2; 3 = exact path) | pair-list base << 2
            unsigned code = 0;
            if (fc >= 3) {
                code = 3;
            } else if (fc > 0) {
                const int base = atomicAdd(&S.npairs, fc);
                if (base + fc <= NP) {
                    S.pij[base] = (unsigned)i | (cand << 16);
                    if (fc == 2) S.pij[base + 1] = (unsigned)i | (cand & 0xFFFF0000u);
                    code = (unsigned)fc | ((unsigned)base << 2);
                } else {
                    // no room: exact path.  The reserved entry below NP (at most one) must still be well formed
                    // for the evaluation loop: a self pair (distance 0) contributes nothing.
                    if (base < NP) S.pij[base] = (unsigned)i | ((unsigned)i << 16);
                    code = 3;
                }
            }
            S.pcode[i] = (unsigned short)code;
        }
        PSIM_TICK(4);
        named_sync<kBarPairs, T>();   // (2) pair list complete
        PSIM_TICK(5);

        // ---- C: exact pair evaluation by the last warp, one lane per pair; meanwhile the other warps bin the NEXT tile's
        //      particles into the other cell table (phase A of the next tile) ----------------------------------------
        if (warp == NW - 1) {
            const int np = min(S.npairs, NP);
#pragma unroll 1
            for (int u = lane; u < np; u += 32) {
                const unsigned ij = S.pij[u];
                PSIM_GUARD((ij & 0xFFFFu) < (unsigned)D::PTOT && (ij >> 16) < (unsigned)D::PTOT, 2, continue)
                const double2 a = st.xy[ij & 0xFFFFu], b = st.xy[ij >> 16];
                const double dx = __dsub_rn(b.x, a.x), dy = __dsub_rn(b.y, a.y);
                const double r2 = pair_r2(dx, dy);
                double cx = 0.0, cy = 0.0;
                if (!(r2 > kCutoff2) && r2 != 0.0) pair_contrib(dx, dy, r2, cx, cy);
                S.pres[u] = make_double2(cx, cy);   // (+0, +0) for a prefilter false positive: neutral in the sum
            }
        } else if (first + (it + 1) * G < P.ntiles) {
            TileWalker nxt = walk;
            nxt.advance(P);
            bin_own(sb ^ 1, (unsigned)(((it + 1) >> 1) & 1), (P.tr_base + nxt.lr) * TS, nxt.tc * TS);
        }
        if (!kStoreAcc) {
            // particles on the exact path (three or more candidates, or no room in the pair list): rare, and the
            // only code with calls -- kept out of the move phase.  The acceleration is folded into the velocity
            // right here (same two roundings as move_particle), the move phase then skips that update.
#pragma unroll 1
            for (int i = tid; i < n_own; i += T) {
                if (S.pcode[i] != 3u) continue;
                // (the cell is recomputed: pcell already holds the next tile's entries)
                const double2 pi = st.xy[i];
                int srow, scol;
                cell_of(pi.x, pi.y, P.bincnt, srow, scol);
                const double2 a = slow_force<W>(st.xy, head, next, i, (srow - (r0 - 1)) * W + (scol - (c0 - 1)));
                double2 v = st.v[i];
                v.x = __dadd_rn(v.x, __dmul_rn(a.x, kDt));
                v.y = __dadd_rn(v.y, __dmul_rn(a.y, kDt));
                st.v[i] = v;
                S.pcode[i] = 4u;   // velocity already advanced
            }
        }
        PSIM_TICK(6);
        named_sync<kBarForces, T + 32>();   // (3) contributions ready, other table clean
        PSIM_TICK(7);
        if (tid == 0) {
            S.hw_pairs = max(S.hw_pairs, S.npairs);
            S.npairs = 0;   // next appended to after barrier 1 of the next tile
            S.prev_lr = lr;
            S.prev_tc = tc;
        }

        // ---- D: sum, move, new cell; E: re-tile (stayers -> other stripe buffer, leavers -> outbox, boundary
        //      cells -> halo lists).  No barrier in between: slots come from warp-aggregated atomics. -----------
#ifdef PSIM_PHASE_TIMERS
        int dbg_class = 0;
#endif
        const size_t gbase = (size_t)(lr * P.ntx + tc) * CAP;
#pragma unroll 1
        for (int i0 = 0; i0 < n_own; i0 += T) {   // uniform trip count: the ballots below need whole warps
            const int i = i0 + tid;
            const bool valid = i < n_own;
            double x = 0.0, y = 0.0, ax = 0.0, ay = 0.0;
            double2 v = make_double2(0.0, 0.0);
            int nrow = 0, ncol = 0, id = 0;
            bool stay = false;
            if (valid) {
                const unsigned code = S.pcode[i], fc = code & 3u;
                PSIM_GUARD(code == 4u || (fc == 3u ? code == 3u : (fc == 0u ? code == 0u : ((code >> 2) + fc) <= (unsigned)NP)), 3)
                const double2 p = st.xy[i];
                v = st.v[i];
                id = st.id[i];
                x = p.x;
                y = p.y;
                if (kStoreAcc && fc == 3u) {
                    // (the cell is recomputed: pcell may already hold the next tile's entries)
                    int srow, scol;
                    cell_of(x, y, P.bincnt, srow, scol);
                    const double2 a = slow_force<W>(st.xy, head, next, i, (srow - (r0 - 1)) * W + (scol - (c0 - 1)));
                    ax = a.x;
                    ay = a.y;
                } else if (fc != 0u) {
                    const double2 c0_ = S.pres[(code >> 2) & 0x3FFFu];
                    ax = __dadd_rn(ax, c0_.x);
                    ay = __dadd_rn(ay, c0_.y);
                    if (fc == 2u) {
                        const double2 c1_ = S.pres[((code >> 2) & 0x3FFFu) + 1];
                        ax = __dadd_rn(ax, c1_.x);
                        ay = __dadd_rn(ay, c1_.y);
                    }
                }
                if (!kStoreAcc && code == 4u) {
                    // exact-path particle: v already holds v + a dt
                    x = __dadd_rn(x, __dmul_rn(v.x, kDt));
                    y = __dadd_rn(y, __dmul_rn(v.y, kDt));
                    reflect_particle(x, y, v.x, v.y, P.size);
                } else
                move_particle(x, y, v.x, v.y, ax, ay, P.size);
                cell_of(x, y, P.bincnt, nrow, ncol);
                stay = (unsigned)(nrow - r0) < (unsigned)TS && (unsigned)(ncol - c0) < (unsigned)TS;
            }
            // destination slots: one atomic per warp, lanes take consecutive slots after the warp's base
            const unsigned sm = __ballot_sync(0xffffffffu, stay), lm = __ballot_sync(0xffffffffu, valid && !stay);
            int sbase = 0, lbase = 0;
            if (lane == 0) {
                if (sm) sbase = atomicAdd(&S.n_stay, __popc(sm));
                if (lm) lbase = atomicAdd(&S.n_leave, __popc(lm));
            }
            sbase = __shfl_sync(0xffffffffu, sbase, 0);
            if (lm) lbase = __shfl_sync(0xffffffffu, lbase, 0);
#ifdef PSIM_PHASE_TIMERS
            {
                const int er_ = nrow - r0, ec_ = ncol - c0;
                const bool corner_ = stay && (er_ == 0 || er_ == TS - 1) && (ec_ == 0 || ec_ == TS - 1);
                dbg_class |= (lm ? 1 : 0) | (__any_sync(0xffffffffu, corner_) ? 2 : 0);
            }
#endif
            PSIM_GUARD(!stay || (unsigned)(sbase + __popc(sm & lt_mask)) < (unsigned)CAP, 4, stay = false)
            if (stay) {
                const double2 q = make_double2(x, y);
                const size_t d = gbase + (size_t)(sbase + __popc(sm & lt_mask));
                P.pos_out[d] = q;
                P.vel_out[d] = v;
                P.id_out[d] = id;
                if (kStoreAcc) P.acc_out[d] = make_double2(ax, ay);
                // boundary cells: append to the edge list(s) and, in a corner cell, the corner list
                // (with TS >= 3 a cell is on at most one of N/S and one of W/E)
                const int er = nrow - r0, ec = ncol - c0;
                const bool n_ = er == 0, s_ = er == TS - 1, w_ = ec == 0, e_ = ec == TS - 1;
                if (n_ | s_ | w_ | e_) {
                    unsigned lists = 0;   // up to three list ids, 4 bits each
                    int nl = 0;
                    if (n_ | s_) { lists = s_ ? 1u : 0u; nl = 1; }
                    if (w_ | e_) { lists |= (e_ ? 3u : 2u) << (4 * nl); ++nl; }
                    if (nl == 2) { lists |= (4u + (s_ ? 2u : 0u) + (e_ ? 1u : 0u)) << 8; nl = 3; }
                    double2* ohxy = reinterpret_cast<double2*>(row_ptr(P.exp_out, P.L, lr) + P.L.off_hxy) + (size_t)tc * D::HL;
                    char* prow = kPeer ? peer_row_of(P, lr) : nullptr;
                    double2* phxy = prow ? reinterpret_cast<double2*>(prow + P.L.off_hxy) + (size_t)tc * D::HL : nullptr;
#pragma unroll 1
                    for (int k = 0; k < nl; ++k) {
                        const int list = (int)((lists >> (4 * k)) & 15u);
                        const int idx = atomicAdd(&S.hout[list], 1);
                        if (idx < halo_cap(list, HE, HC)) {
                            ohxy[halo_offset(list, HE, HC) + idx] = q;
                            if (phxy) phxy[halo_offset(list, HE, HC) + idx] = q;
                        }
                    }
                }
            } else if (valid) {
                const int rank = lbase + __popc(lm & lt_mask);
                if (rank < CO) {
                    OutRec* oobox = reinterpret_cast<OutRec*>(row_ptr(P.exp_out, P.L, lr) + P.L.off_obox) + (size_t)tc * CO;
                    OutRec rec;
                    rec.x = x; rec.y = y; rec.vx = v.x; rec.vy = v.y;
                    rec.ax = kStoreAcc ? ax : 0.0;
                    rec.ay = kStoreAcc ? ay : 0.0;
                    rec.id = id;
                    rec.row = nrow; rec.col = ncol; rec.pad = 0;
                    oobox[rank] = rec;
                    if (kPeer) {
                        if (char* prow = peer_row_of(P, lr)) (reinterpret_cast<OutRec*>(prow + P.L.off_obox) + (size_t)tc * CO)[rank] = rec;
                    }
                }
                // a particle may not skip a whole tile in one step
                if ((unsigned)(nrow - r0 + TS) >= (unsigned)(3 * TS) || (unsigned)(ncol - c0 + TS) >= (unsigned)(3 * TS))
                    atomicOr(&S.flags, kErrLostParticle);
            }
        }
        walk.advance(P);
        PSIM_TICK(8);
        PSIM_HIST(0, 2);
        PSIM_HIST(1, 4);
        PSIM_HIST_KEEP(4 + (dbg_class & 3), 8);
        PSIM_HIST(2, 8);
        PSIM_HIST(3, 1);
    }
#ifdef PSIM_PHASE_TIMERS
    if (lane == 0) for (int k = 0; k < 9; ++k) atomicAdd(&g_phase_cycles[warp][k], (unsigned long long)tacc[k]);
#endif
    named_sync<kBarPairs, T>();   // all appends and slot reservations of the last tile are done
    if (tid < 9) publish_counts();
    named_sync<kBarPairs, T>();
    if (tid == 0) {
        if (S.flags) atomicOr(P.err, S.flags);
        atomicMax(P.err + 1, S.hw_leave);
        atomicMax(P.err + 2, S.hw_halo);
        atomicMax(P.err + 3, S.hw_tile);
        atomicMax(P.err + 4, S.hw_apron);
        atomicMax(P.err + 7, S.hw_pairs);
    }
}

// ------------------------------------------------------------------------------------------
// initial tiling
// ------------------------------------------------------------------------------------------
// one thread per input record: owned rows only; arrival order inside a tile is arbitrary, which is
// harmless because forces are summed in a canonical order.
__global__ void __launch_bounds__(256) tile_fill_kernel(const particle_t* __restrict__ p, int n, int id0, int bincnt,
                                                        int ts, int cap, int ntx, int tr_begin, int tr_end,
                                                        int tr_base, double2* __restrict__ pos, double2* __restrict__ vel,
                                                        int* __restrict__ sid, int* __restrict__ tcount,
                                                        int* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2* q = reinterpret_cast<const double2*>(p + i);
    const double2 a = q[0], b = q[1];
    const int tr = axis_cell(a.x, bincnt) / ts, tc = axis_cell(a.y, bincnt) / ts;
    if (tr < tr_begin || tr >= tr_end) return;
    const int lt = (tr - tr_base) * ntx + tc;
    const int slot = atomicAdd(tcount + lt, 1);
    if (slot >= cap) {
        atomicOr(err, kErrTileOverflow);
        return;
    }
    const size_t d = (size_t)lt * cap + slot;
    pos[d] = a;
    vel[d] = b;
    sid[d] = id0 + i;
}

// first export of halo lists (parity 0) from the freshly filled tiles; outboxes start empty.
template <int TS>
__global__ void __launch_bounds__(128) tile_export_kernel(const TileParams P) {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    constexpr int CAP = C::CAP, HE = C::HE, HC = C::HC;
    __shared__ int s_hout[8];
    const int tid = threadIdx.x;
    const TileCoord c = tile_coord(P, blockIdx.x);
    const int r0 = c.tr * TS, c0 = c.tc * TS;
    const size_t gbase = (size_t)c.lt * CAP;
    const int n = min(P.tcount[c.lt], CAP);
    if (tid < 8) s_hout[tid] = 0;
    __syncthreads();
    char* orow = row_ptr(P.exp_out, P.L, c.lr);
    double2* ohxy = reinterpret_cast<double2*>(orow + P.L.off_hxy) + (size_t)c.tc * D::HL;
    for (int i = tid; i < n; i += blockDim.x) {
        const double2 q = P.pos_in[gbase + i];
        const int er = axis_cell(q.x, P.bincnt) - r0, ec = axis_cell(q.y, P.bincnt) - c0;
        const bool n_ = er == 0, s_ = er == TS - 1, w_ = ec == 0, e_ = ec == TS - 1;
        if (n_ | s_ | w_ | e_) {
            auto put = [&](int list) {
                const int idx = atomicAdd(&s_hout[list], 1);
                if (idx < halo_cap(list, HE, HC)) ohxy[halo_offset(list, HE, HC) + idx] = q;
            };
            if (n_) put(0);
            if (s_) put(1);
            if (w_) put(2);
            if (e_) put(3);
            if (n_ && w_) put(4);
            if (n_ && e_) put(5);
            if (s_ && w_) put(6);
            if (s_ && e_) put(7);
        }
    }
    __syncthreads();
    if (tid < 9) {
        int* ec = reinterpret_cast<int*>(orow + P.L.off_cnt) + (size_t)c.tc * 16;
        if (tid < 8) {
            const int k = s_hout[tid], cap = halo_cap(tid, HE, HC);
            if (k > cap) atomicOr(P.err, kErrHaloOverflow);
            ec[tid] = min(k, cap);
        } else {
            ec[8] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// observation: gather the owned particles (tile stripes + outbox records that land in owned rows)
// into a compact SoA.  One CTA per tile (ghost rows included: their outboxes may hold particles that
// have just crossed into this slab).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tile_gather_kernel(const TileParams P, const double2* __restrict__ acc, int ts, int cap,
                                                          int co, int tr_begin, int tr_end, bool have_acc,
                                                          double* __restrict__ gx, double* __restrict__ gy,
                                                          double* __restrict__ gvx, double* __restrict__ gvy,
                                                          double* __restrict__ gax, double* __restrict__ gay,
                                                          int* __restrict__ gid, int* __restrict__ cursor) {
    __shared__ int s_base, s_take[64], s_ntake;
    const int lr = blockIdx.x / P.ntx, tc = blockIdx.x % P.ntx;
    const int tr = P.tr_base + lr;
    const int lt = lr * P.ntx + tc;
    const bool owned = tr >= tr_begin && tr < tr_end;
    const int n_tile = owned ? min(P.tcount[lt], cap) : 0;
    const char* row = row_ptr(P.exp_in, P.L, lr);
    const bool row_valid = tr >= 0 && tr < P.nty;
    const int n_out = row_valid ? min((reinterpret_cast<const int*>(row + P.L.off_cnt) + (size_t)tc * 16)[8], co) : 0;
    const OutRec* ob = reinterpret_cast<const OutRec*>(row + P.L.off_obox) + (size_t)tc * co;
    if (threadIdx.x == 0) {
        int k = 0;
        for (int e = 0; e < n_out && k < 64; ++e) {
            const int dtr = ob[e].row / ts;
            if (dtr >= tr_begin && dtr < tr_end) s_take[k++] = e;
        }
        s_ntake = k;
        s_base = (n_tile + k) ? atomicAdd(cursor, n_tile + k) : 0;
    }
    __syncthreads();
    const int base = s_base;
    const size_t gbase = (size_t)lt * cap;
    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
        const double2 p = P.pos_in[gbase + i], v = P.vel_in[gbase + i];
        const double2 a = have_acc ? acc[gbase + i] : make_double2(0.0, 0.0);
        gx[base + i] = p.x;
        gy[base + i] = p.y;
        gvx[base + i] = v.x;
        gvy[base + i] = v.y;
        gax[base + i] = a.x;
        gay[base + i] = a.y;
        gid[base + i] = P.id_in[gbase + i];
    }
    for (int k = threadIdx.x; k < s_ntake; k += blockDim.x) {
        const OutRec r = ob[s_take[k]];
        const int d = base + n_tile + k;
        gx[d] = r.x;
        gy[d] = r.y;
        gvx[d] = r.vx;
        gvy[d] = r.vy;
        gax[d] = have_acc ? r.ax : 0.0;
        gay[d] = have_acc ? r.ay : 0.0;
        gid[d] = r.id;
    }
}

// Write-back in ORIGINAL particle order straight from the tiles (the drivers' view of `parts`, reference
// part1/main.cpp:135-136 / part3/main.cu:134-136): out[id] = {x y vx vy ax ay}, or only xy[id] = {x y}.  One CTA per
// tile; particles still sitting in an outbox (they left their tile in the last step) are written by the tile that owns
// the outbox.  Every owned particle is in exactly one stripe or one outbox, so no intermediate compaction is needed.
__global__ void __launch_bounds__(128) tile_writeback_kernel(const TileParams P, const double2* __restrict__ acc, int cap, int co,
                                                             int tr_begin, int tr_end, int ts, bool have_acc,
                                                             particle_t* __restrict__ out, double2* __restrict__ out_xy) {
    const int lr = blockIdx.x / P.ntx, tc = blockIdx.x % P.ntx;
    const int tr = P.tr_base + lr;
    const int lt = lr * P.ntx + tc;
    const bool owned = tr >= tr_begin && tr < tr_end;
    const int n_tile = owned ? min(P.tcount[lt], cap) : 0;
    const size_t gbase = (size_t)lt * cap;
    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
        const int id = P.id_in[gbase + i];
        const double2 p = P.pos_in[gbase + i];
        if (out_xy) {
            out_xy[id] = p;
        } else {
            double2* q = reinterpret_cast<double2*>(out + id);
            q[0] = p;
            q[1] = P.vel_in[gbase + i];
            q[2] = have_acc ? acc[gbase + i] : make_double2(0.0, 0.0);
        }
    }
    const bool row_valid = tr >= 0 && tr < P.nty;
    const char* row = row_ptr(P.exp_in, P.L, lr);
    const int n_out = row_valid ? min((reinterpret_cast<const int*>(row + P.L.off_cnt) + (size_t)tc * 16)[8], co) : 0;
    const OutRec* ob = reinterpret_cast<const OutRec*>(row + P.L.off_obox) + (size_t)tc * co;
    for (int k = threadIdx.x; k < n_out; k += blockDim.x) {
        const OutRec r = ob[k];
        const int dtr = r.row / ts;
        if (dtr < tr_begin || dtr >= tr_end) continue;   // migrated to another slab
        if (out_xy) {
            out_xy[r.id] = make_double2(r.x, r.y);
        } else {
            double2* q = reinterpret_cast<double2*>(out + r.id);
            q[0] = make_double2(r.x, r.y);
            q[1] = make_double2(r.vx, r.vy);
            q[2] = have_acc ? make_double2(r.ax, r.ay) : make_double2(0.0, 0.0);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct TiledEngine {
    DeviceArena mem;
    int ts = 0, cap = 0, he = 0, hc = 0, co = 0, hl = 0, threads = 0, ctas_per_sm = 1;
    size_t smem = 0;
    int sms = 148;
    int ntx = 0;                   // tiles per side
    int tr_begin = 0, tr_end = 0;  // owned tile rows (global)
    int lrows = 0;                 // owned rows
    int lrows_alloc = 0;           // owned + 2 ghost rows
    ExportLayout L{};
    // stripes and exports, double buffered by step parity: buffer [parity] is what the next step reads
    double2 *pos[2] = {nullptr, nullptr}, *vel[2] = {nullptr, nullptr}, *acc = nullptr;
    int* sid[2] = {nullptr, nullptr};
    int* tcount = nullptr;
    char* exports[2] = {nullptr, nullptr};
    size_t export_bytes = 0;  // one parity
    int parity = 0;
    bool acc_valid = false;    // acc holds the accelerations of the step that produced buffer [parity]
    bool ghost_fresh = false;  // ghost rows hold the neighbours' exports of the current parity
    int comm_reserve_ctas = 16;  // CTA slots left free for the exchange kernels in slab mode (PSIM_COMM_RESERVE)
    // gather scratch
    DeviceArena gmem;
    SoAView g{};
    int* g_cursor = nullptr;
    int g_capacity = 0;
};

void tiled_slab_rows(int ntx, int rank, int nranks, int* begin, int* end);

static ExportLayout make_layout(int ntx, int hl, int co) {
    ExportLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    L.off_cnt = take((size_t)ntx * 16 * sizeof(int));
    L.off_hxy = take((size_t)ntx * hl * sizeof(double2));
    L.off_obox = take((size_t)ntx * co * sizeof(OutRec));
    L.row_bytes = off;
    return L;
}

static TileParams make_params(psim_sim* sim, TiledEngine* e, int parity_in) {
    TileParams P{};
    P.pos_in = e->pos[parity_in]; P.vel_in = e->vel[parity_in]; P.id_in = e->sid[parity_in];
    P.pos_out = e->pos[parity_in ^ 1]; P.vel_out = e->vel[parity_in ^ 1]; P.id_out = e->sid[parity_in ^ 1];
    P.acc_out = e->acc;
    P.tcount = e->tcount;
    P.exp_in = e->exports[parity_in];
    P.exp_out = e->exports[parity_in ^ 1];
    P.L = e->L;
    P.ntx = e->ntx;
    P.nty = e->ntx;
    P.tr_base = e->tr_begin - 1;
    P.lrow0 = 1;
    P.row_stride = 1;
    P.ntiles = e->lrows * e->ntx;
    P.bincnt = sim->bincnt;
    P.size = sim->size;
    P.err = sim->d_err;
    P.last_lrow = e->lrows;
    P.peer_row[0] = P.peer_row[1] = nullptr;
    if (sim->p2p) {
        const int po = parity_in ^ 1;
        if (sim->rank > 0) {
            int b, en;
            tiled_slab_rows(e->ntx, sim->rank - 1, sim->nranks, &b, &en);
            P.peer_row[0] = sim->peer_exports[0][po] + e->L.row_bytes * (size_t)(en - b + 1);   // its upper ghost row
        }
        if (sim->rank < sim->nranks - 1) P.peer_row[1] = sim->peer_exports[1][po];              // its lower ghost row (row 0)
    }
    return P;
}

template <int TS>
static int launch_step(psim_sim* sim, TiledEngine* e, int parity_in, bool store_acc, int lrow0, int nrows, int row_stride,
                       bool allow_peer, cudaStream_t s) {
    if (nrows <= 0) return PSIM_OK;
    TileParams P = make_params(sim, e, parity_in);
    if (!allow_peer) P.peer_row[0] = P.peer_row[1] = nullptr;
    P.lrow0 = lrow0;
    P.row_stride = row_stride;
    P.ntiles = nrows * e->ntx;
    // slabs: leave a few CTA slots free so that the exchange kernels (NCCL send / recv on the high-priority stream) can
    // become resident next to the persistent interior-row CTAs instead of waiting for them to finish
    const int reserve = (sim->nranks > 1 && !sim->p2p) ? e->comm_reserve_ctas : 0;
    const int grid = std::min(P.ntiles, std::max(e->sms, e->sms * e->ctas_per_sm - reserve));
    const bool peer = P.peer_row[0] || P.peer_row[1];
    constexpr int kThreads = TileCfg<TS>::THREADS + 32;
    constexpr size_t kSmem = sizeof(TileSmem<TS>);
    if (store_acc) {
        if (peer) tile_step_kernel<TS, true, true><<<grid, kThreads, kSmem, s>>>(P);
        else tile_step_kernel<TS, true, false><<<grid, kThreads, kSmem, s>>>(P);
    } else {
        if (peer) tile_step_kernel<TS, false, true><<<grid, kThreads, kSmem, s>>>(P);
        else tile_step_kernel<TS, false, false><<<grid, kThreads, kSmem, s>>>(P);
    }
    ++sim->launches;
    return PSIM_OK;
}

static int launch_step_ts(psim_sim* sim, TiledEngine* e, int parity_in, bool store_acc, int lrow0, int nrows, cudaStream_t s,
                          int row_stride = 1, bool allow_peer = true) {
    switch (e->ts) {
        case 16: return launch_step<16>(sim, e, parity_in, store_acc, lrow0, nrows, row_stride, allow_peer, s);
        case 32: return launch_step<32>(sim, e, parity_in, store_acc, lrow0, nrows, row_stride, allow_peer, s);
        case 64: return launch_step<64>(sim, e, parity_in, store_acc, lrow0, nrows, row_stride, allow_peer, s);
    }
    return fail(PSIM_ERR_INVALID, "tile size %d not instantiated", e->ts);
}

template <int TS>
static int configure(TiledEngine* e) {
    e->cap = TileCfg<TS>::CAP;
    e->he = TileCfg<TS>::HE;
    e->hc = TileCfg<TS>::HC;
    e->co = TileCfg<TS>::CO;
    e->hl = TileDims<TS>::HL;
    e->threads = TileCfg<TS>::THREADS + 32;
    e->smem = sizeof(TileSmem<TS>);
    auto prepare = [&](auto kernel) -> int {
        PSIM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
        PSIM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        return PSIM_OK;
    };
    PSIM_TRY(prepare(tile_step_kernel<TS, true, true>));
    PSIM_TRY(prepare(tile_step_kernel<TS, true, false>));
    PSIM_TRY(prepare(tile_step_kernel<TS, false, true>));
    PSIM_TRY(prepare(tile_step_kernel<TS, false, false>));
    int per_sm = 0;
    PSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_step_kernel<TS, false, false>, TileCfg<TS>::THREADS + 32, e->smem));
    e->ctas_per_sm = std::max(1, per_sm);
    if (const char* r = std::getenv("PSIM_COMM_RESERVE")) e->comm_reserve_ctas = std::max(0, std::atoi(r));
    if (const char* cap = std::getenv("PSIM_CTAS_PER_SM")) {   // tuning / profiling knob
        const int c = std::atoi(cap);
        if (c >= 1) e->ctas_per_sm = std::min(e->ctas_per_sm, c);
    }
    return PSIM_OK;
}

int tiled_tile_rows(int bincnt, int ts) { return (bincnt + ts - 1) / ts; }

void tiled_slab_rows(int ntx, int rank, int nranks, int* begin, int* end) {
    // contiguous tile rows, remainder spread over the first ranks
    const int q = ntx / nranks, r = ntx % nranks;
    *begin = rank * q + std::min(rank, r);
    *end = *begin + q + (rank < r ? 1 : 0);
}

int tiled_default_tile(int bincnt) {
    // 32-cell tiles amortise the apron and the per-tile bookkeeping best; small boxes keep 16-cell tiles so that there
    // are enough tiles to spread over the SMs (measured: 100 k particles, 23^2 tiles of 32 cells: 17.5 us/step vs 19.6)
    const long long ntx32 = (bincnt + 31) / 32;
    return ntx32 * ntx32 >= 256 ? 32 : 16;
}

int tiled_create(psim_sim* sim, const psim_config* cfg, const particle_t* parts, int n, bool parts_on_device,
                 bool* unsuitable) {
    *unsuitable = false;
    cudaStream_t s = sim->stream;
    int ts = cfg->tile_cells;
    if (ts == 0) ts = tiled_default_tile(sim->bincnt);
    if (ts != 16 && ts != 32 && ts != 64) return fail(PSIM_ERR_INVALID, "tile_cells must be 16, 32 or 64 (got %d)", ts);
    auto* e = new TiledEngine();
    sim->tiled = e;
    e->ts = ts;
    PSIM_CUDA(cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, sim->device));
    if (ts == 16) PSIM_TRY(configure<16>(e));
    if (ts == 32) PSIM_TRY(configure<32>(e));
    if (ts == 64) PSIM_TRY(configure<64>(e));
    e->ntx = tiled_tile_rows(sim->bincnt, ts);
    if (sim->nranks > e->ntx) return fail(PSIM_ERR_INVALID, "more slabs (%d) than tile rows (%d)", sim->nranks, e->ntx);
    tiled_slab_rows(e->ntx, sim->rank, sim->nranks, &e->tr_begin, &e->tr_end);
    e->lrows = e->tr_end - e->tr_begin;
    e->lrows_alloc = e->lrows + 2;
    sim->row_begin = e->tr_begin * ts;
    sim->row_end = std::min(e->tr_end * ts, sim->bincnt);
    e->L = make_layout(e->ntx, e->hl, e->co);
    e->export_bytes = e->L.row_bytes * (size_t)e->lrows_alloc;

    const size_t slots = (size_t)e->lrows_alloc * e->ntx * e->cap;
    // the export buffers are separate allocations (their CUDA IPC handles are shared with the neighbour slabs) ...
    for (int b = 0; b < 2; ++b) {
        PSIM_TRY(e->mem.alloc(&e->exports[b], e->export_bytes));
        PSIM_CUDA(cudaMemsetAsync(e->exports[b], 0, e->export_bytes, s));
    }
    // ... everything else comes out of one block
    PSIM_TRY(e->mem.reserve(slots * (2 * (2 * sizeof(double2) + sizeof(int)) + sizeof(double2)) +
                            sizeof(int) * (size_t)e->lrows_alloc * e->ntx + 16 * 256));
    for (int b = 0; b < 2; ++b) {
        PSIM_TRY(e->mem.alloc(&e->pos[b], slots));
        PSIM_TRY(e->mem.alloc(&e->vel[b], slots));
        PSIM_TRY(e->mem.alloc(&e->sid[b], slots));
    }
    PSIM_TRY(e->mem.alloc(&e->acc, slots));
    PSIM_TRY(e->mem.alloc(&e->tcount, (size_t)e->lrows_alloc * e->ntx));
    PSIM_CUDA(cudaMemsetAsync(e->tcount, 0, sizeof(int) * (size_t)e->lrows_alloc * e->ntx, s));
    PSIM_CUDA(cudaMemsetAsync(e->acc, 0, sizeof(double2) * slots, s));
    // fill: device input is read in place; host input is streamed through a bounded staging buffer
    // (a slab keeps only its own rows, so no rank ever holds the whole array on the device)
    {
        DeviceArena stage;
        particle_t* d_stage = nullptr;
        const int chunk = parts_on_device ? n : std::min(n, 4 << 20);
        if (!parts_on_device && n > 0) PSIM_TRY(stage.alloc(&d_stage, (size_t)chunk));
        for (int off = 0; off < n; off += chunk) {
            const int m = std::min(chunk, n - off);
            const particle_t* src = parts + off;
            if (!parts_on_device) {
                PSIM_CUDA(cudaMemcpyAsync(d_stage, parts + off, sizeof(particle_t) * (size_t)m, cudaMemcpyHostToDevice, s));
                src = d_stage;
            }
            tile_fill_kernel<<<(m + 255) / 256, 256, 0, s>>>(src, m, off, sim->bincnt, ts, e->cap, e->ntx, e->tr_begin,
                                                             e->tr_end, e->tr_begin - 1, e->pos[0], e->vel[0], e->sid[0],
                                                             e->tcount, sim->d_err);
            ++sim->launches;   // (no host synchronisation per chunk: copy and kernel are ordered by the stream)
        }
        PSIM_CUDA(cudaGetLastError());
        if (!parts_on_device) PSIM_CUDA(cudaStreamSynchronize(s));   // the staging buffer is freed below
        stage.release();
    }
    // suitability: the densest tile must leave headroom for fluctuations, else the caller falls back
    {
        std::vector<int> h((size_t)e->lrows_alloc * e->ntx);
        PSIM_CUDA(cudaMemcpyAsync(h.data(), e->tcount, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, s));
        PSIM_CUDA(cudaStreamSynchronize(s));
        int worst = 0;
        for (int v : h) worst = std::max(worst, v);
        if (worst + worst / 4 + 16 > e->cap) {
            *unsuitable = true;
            PSIM_CUDA(cudaMemsetAsync(sim->d_err, 0, sizeof(int), s));
            return fail(PSIM_ERR_UNSUPPORTED,
                        "tiled engine: densest %dx%d-cell tile holds %d particles, capacity %d leaves no headroom", ts, ts,
                        worst, e->cap);
        }
    }
    TileParams P = make_params(sim, e, 0);   // reads stripes [0] ...
    P.exp_out = e->exports[0];               // ... and writes the exports of the same parity
    const int grid = e->lrows * e->ntx;
    if (ts == 16) tile_export_kernel<16><<<grid, 128, 0, s>>>(P);
    if (ts == 32) tile_export_kernel<32><<<grid, 128, 0, s>>>(P);
    if (ts == 64) tile_export_kernel<64><<<grid, 128, 0, s>>>(P);
    ++sim->launches;
    PSIM_CUDA(cudaGetLastError());
    e->parity = 0;
    e->acc_valid = true;  // zeros
    return PSIM_OK;
}

int tiled_step(psim_sim* sim, int nsteps, int flags) {
    TiledEngine* e = sim->tiled;
    cudaStream_t s = sim->stream;
    const bool slabs = sim->nranks > 1;
    if (slabs && !sim->comm) return fail(PSIM_ERR_STATE, "slab %d/%d is not connected: call psim_comm_connect first", sim->rank, sim->nranks);
    for (int step = 0; step < nsteps; ++step) {
        const bool store = (flags & PSIM_STEP_ACCEL_ALL) || (!(flags & PSIM_STEP_ACCEL_NONE) && step == nsteps - 1);
        if (!slabs) {
            PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 1, e->lrows, s));
        } else if (sim->p2p) {
            // Slab step, peer-memory flavour: the first and last owned tile rows go first, and that launch stores their
            // exports (halo lists + migrants) into the local buffers AND straight into the neighbours' ghost rows over
            // NVLink; a flag then tells the neighbours that my boundary rows of this step are done, and the interior
            // rows follow.  My boundary rows may start once both neighbours have flagged the previous step: their
            // stores into my ghost rows are complete and they no longer read the ghost rows I am about to overwrite.
            // No exchange kernel, no copy: the transfer is fused into the step kernel.
            if (!e->ghost_fresh) {   // first step after create: the neighbours' initial exports travel once through NCCL
                PSIM_TRY(tiled_exchange(sim, e->parity, s));
                e->ghost_fresh = true;
            }
            // The boundary launch (stream B, high priority) and the interior launch (stream I = the handle's stream) of one
            // step touch disjoint tiles and run CONCURRENTLY; each depends only on BOTH launches of the previous step:
            //   boundary(s) after interior(s-1) [its rows 2 / lrows-1 exports], boundary(s-1), the neighbours' flags
            //   interior(s) after boundary(s-1) [rows 1 / lrows exports], interior(s-1)
            cudaStream_t sb = sim->comm_stream;
            const unsigned k = sim->p2p_steps & 1u, kp = k ^ 1u;   // event slots of this step / the previous one
            if (sim->p2p_steps == 0) {   // first step: everything enqueued so far on the handle's stream (create, initial exchange)
                PSIM_CUDA(cudaEventRecord(sim->ev_i[kp], s));
                PSIM_CUDA(cudaEventRecord(sim->ev_b[kp], s));
            }
            PSIM_CUDA(cudaStreamWaitEvent(sb, sim->ev_i[kp], 0));
            PSIM_TRY(comm_p2p_wait(sim, sb));
            if (e->lrows <= 2) PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 1, e->lrows, sb));
            else PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 1, 2, sb, e->lrows - 1));   // rows 1 and lrows, storing to the peers
            PSIM_CUDA(cudaEventRecord(sim->ev_b[k], sb));
            PSIM_TRY(comm_p2p_signal(sim, sb));
            PSIM_CUDA(cudaStreamWaitEvent(s, sim->ev_b[kp], 0));
            if (e->lrows > 2) PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 2, e->lrows - 2, s, 1, /*peer=*/false));
            PSIM_CUDA(cudaEventRecord(sim->ev_i[k], s));
            if (step == nsteps - 1) PSIM_CUDA(cudaStreamWaitEvent(s, sim->ev_b[k], 0));   // the handle's stream covers the whole batch
        } else {
            // Slab step, NCCL flavour: the first and last owned tile rows go first; as soon as they are
            // done their exports (halo lists + migrants) travel to the neighbours on the exchange stream while the
            // interior rows are computed.  The next step's boundary rows wait for both.
            if (!e->ghost_fresh) {   // first step after create: the neighbours' initial exports
                PSIM_TRY(tiled_exchange(sim, e->parity, s));
                e->ghost_fresh = true;
            }
            if (e->lrows <= 2) PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 1, e->lrows, s));
            else PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 1, 2, s, e->lrows - 1));   // rows 1 and lrows in one launch
            PSIM_CUDA(cudaEventRecord(sim->ev_boundary, s));
            PSIM_CUDA(cudaStreamWaitEvent(sim->comm_stream, sim->ev_boundary, 0));
            PSIM_TRY(tiled_exchange(sim, e->parity ^ 1, sim->comm_stream));
            PSIM_CUDA(cudaEventRecord(sim->ev_exchanged, sim->comm_stream));
            if (e->lrows > 2) PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 2, e->lrows - 2, s));
            PSIM_CUDA(cudaStreamWaitEvent(s, sim->ev_exchanged, 0));   // the next step (or an observation) sees fresh ghost rows
        }
        e->parity ^= 1;
        e->acc_valid = store;
        ++sim->steps_done;
    }
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

int tiled_view(psim_sim* sim, SoAView* out) {
    TiledEngine* e = sim->tiled;
    cudaStream_t s = sim->stream;
    if (!e->g_cursor || e->g_capacity < sim->n_total) {
        e->gmem.release();
        const size_t c = (size_t)sim->n_total + 2;
        PSIM_TRY(e->gmem.alloc(&e->g.x, c));
        PSIM_TRY(e->gmem.alloc(&e->g.y, c));
        PSIM_TRY(e->gmem.alloc(&e->g.vx, c));
        PSIM_TRY(e->gmem.alloc(&e->g.vy, c));
        PSIM_TRY(e->gmem.alloc(&e->g.ax, c));
        PSIM_TRY(e->gmem.alloc(&e->g.ay, c));
        PSIM_TRY(e->gmem.alloc(&e->g.id, c));
        PSIM_TRY(e->gmem.alloc(&e->g_cursor, 1));
        e->g_capacity = sim->n_total;
    }
    PSIM_CUDA(cudaMemsetAsync(e->g_cursor, 0, sizeof(int), s));
    if (sim->nranks > 1 && !e->ghost_fresh) {
        PSIM_TRY(tiled_exchange(sim, e->parity, s));
        e->ghost_fresh = true;
    }
    if (sim->p2p) PSIM_TRY(comm_p2p_wait(sim, s));   // the neighbours' stores into my ghost rows are complete
    TileParams P = make_params(sim, e, e->parity);
    tile_gather_kernel<<<e->lrows_alloc * e->ntx, 128, 0, s>>>(P, e->acc, e->ts, e->cap, e->co, e->tr_begin, e->tr_end,
                                                               e->acc_valid, e->g.x, e->g.y, e->g.vx, e->g.vy, e->g.ax,
                                                               e->g.ay, e->g.id, e->g_cursor);
    ++sim->launches;
    PSIM_CUDA(cudaGetLastError());
    int n = 0;
    PSIM_CUDA(cudaMemcpyAsync(&n, e->g_cursor, sizeof(int), cudaMemcpyDeviceToHost, s));
    PSIM_CUDA(cudaStreamSynchronize(s));
    *out = e->g;
    out->n = n;
    return PSIM_OK;
}

// original-order write-back into a DEVICE buffer of n_total records (exactly one of d_out / d_xy); enqueued on the handle's stream
int tiled_writeback(psim_sim* sim, particle_t* d_out, double2* d_xy) {
    TiledEngine* e = sim->tiled;
    cudaStream_t s = sim->stream;
    if (sim->nranks > 1 && !e->ghost_fresh) {
        PSIM_TRY(tiled_exchange(sim, e->parity, s));
        e->ghost_fresh = true;
    }
    if (sim->p2p) PSIM_TRY(comm_p2p_wait(sim, s));
    TileParams P = make_params(sim, e, e->parity);
    tile_writeback_kernel<<<e->lrows_alloc * e->ntx, 128, 0, s>>>(P, e->acc, e->cap, e->co, e->tr_begin, e->tr_end, e->ts,
                                                                  e->acc_valid, d_out, d_xy);
    ++sim->launches;
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

void tiled_destroy(psim_sim* sim) {
    TiledEngine* e = sim->tiled;
    if (!e) return;
#ifdef PSIM_PHASE_TIMERS
    {
        unsigned long long h[34][12];
        cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof h);
        const double tiles = (double)e->lrows * e->ntx * (double)std::max<long long>(sim->steps_done, 1);
        const char* names[9] = {"top", "wait_full", "A", "bar1", "B", "bar2", "C", "bar3", "DE"};
        const char* pnames[9] = {"top", "bar1", "issue", "bar3", "wait+ingest", "-", "-", "-", "-"};
        for (int w = 0; w < 34; ++w) {
            if (w == 33) for (int k = 0; k < 9; ++k) names[k] = pnames[k];
            if (h[w][0] + h[w][2] == 0) continue;
            std::fprintf(stderr, "[phase cycles/tile] %s%d:", w == 33 ? "producer" : "warp", w);
            double tot = 0;
            for (int k = 0; k < 9; ++k) {
                std::fprintf(stderr, " %s=%.0f", names[k], h[w][k] / tiles);
                tot += h[w][k] / tiles;
            }
            std::fprintf(stderr, " total=%.0f\n", tot);
        }
        unsigned hh[8][64];
        cudaMemcpyFromSymbol(hh, g_hist, sizeof hh);
        const char* hn[8] = {"A", "B", "DE", "wait_full", "DE plain", "DE leaver", "DE corner", "DE leaver+corner"};
        for (int a = 0; a < 8; ++a) {
            std::fprintf(stderr, "[hist %s, 256-cycle bins]", hn[a]);
            for (int b = 0; b < 64; ++b) if (hh[a][b]) std::fprintf(stderr, " %d:%u", b, hh[a][b]);
            std::fprintf(stderr, "\n");
        }
        unsigned long long z[34][12] = {};
        cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z);
    }
#endif
    e->gmem.release();
    e->mem.release();
    delete e;
    sim->tiled = nullptr;
}

long long tiled_bytes(psim_sim* sim) { return sim->tiled ? (long long)(sim->tiled->mem.bytes + sim->tiled->gmem.bytes) : 0; }

void tiled_info(psim_sim* sim, psim_info_t* out) {
    TiledEngine* e = sim->tiled;
    out->tile_cells = e->ts;
    out->tiles_per_side = e->ntx;
    out->tile_capacity = e->cap;
    out->outbox_capacity = e->co;
    out->halo_list_capacity = e->he;
}

}  // namespace psim

extern "C" int psim_slab_rows(int bin_count, int tile_cells, int rank, int nranks, int* row_begin, int* row_end) {
    using namespace psim;
    if (bin_count < 1 || nranks < 1 || rank < 0 || rank >= nranks || !row_begin || !row_end)
        return fail(PSIM_ERR_INVALID, "psim_slab_rows: bad argument");
    const int ts = tile_cells ? tile_cells : kstep_default_tile(bin_count);   // 0: what the default (kstep) engine picks for this box
    if (!kstep_tile_supported(ts)) return fail(PSIM_ERR_INVALID, "tile_cells %d is not a tile size of this build", ts);
    const int ntx = tiled_tile_rows(bin_count, ts);
    if (nranks > ntx) return fail(PSIM_ERR_INVALID, "more slabs (%d) than tile rows (%d)", nranks, ntx);
    int b, e;
    tiled_slab_rows(ntx, rank, nranks, &b, &e);
    *row_begin = b * ts;
    *row_end = std::min(e * ts, bin_count);
    return PSIM_OK;
}

namespace psim {

// accessors for psim_comm.cpp
void tiled_export_buffers(psim_sim* sim, char** parity0, char** parity1, size_t* bytes, size_t* row_bytes, int* lrows, int* ntx) {
    TiledEngine* e = sim->tiled;
    *parity0 = e->exports[0];
    *parity1 = e->exports[1];
    *bytes = e->export_bytes;
    *row_bytes = e->L.row_bytes;
    *lrows = e->lrows;
    *ntx = e->ntx;
}

// one thread: publish my finished step count in the neighbours' flag words (peer memory); the preceding kernel's
// stores are already visible system-wide at its completion, the fence orders the flag behind them
__global__ void flag_store_kernel(int* flag_a, int* flag_b, int value) {
    __threadfence_system();
    if (flag_a) *reinterpret_cast<volatile int*>(flag_a) = value;
    if (flag_b) *reinterpret_cast<volatile int*>(flag_b) = value;
}
void launch_flag_store(int* flag_a, int* flag_b, int value, cudaStream_t s) { flag_store_kernel<<<1, 1, 0, s>>>(flag_a, flag_b, value); }

void tiled_boundary_rows(psim_sim* sim, int parity, char** first_owned, char** last_owned, char** ghost_lo, char** ghost_hi,
                         size_t* row_bytes) {
    TiledEngine* e = sim->tiled;
    char* base = e->exports[parity];
    *ghost_lo = base;
    *first_owned = base + e->L.row_bytes;
    *last_owned = base + e->L.row_bytes * (size_t)e->lrows;
    *ghost_hi = base + e->L.row_bytes * (size_t)(e->lrows + 1);
    *row_bytes = e->L.row_bytes;
}

}  // namespace psim
