// psim_kstep.cu -- the "kstep" engine: tile-resident particles, K time steps fused per launch in shared memory.
//
// Why.  The per-step hot path (bin -> 3x3 force gather -> move, reference part1/serial.cpp:119-131,
// part3/gpu.cu:187-208) moves 80 bytes per particle-step through HBM but needs a few hundred instructions per
// particle to find the ~0.1 in-range pairs; on a B200 the instruction issue rate, not HBM, is what binds a
// one-step-per-launch kernel (ncu: profiles/r1k_*).  Particles move at most ~0.15 cutoff cells per step, so a
// tile that is loaded together with an H-cell halo can be advanced K steps without looking at HBM again:
//   * the gather / re-tile / export machinery runs once per K steps instead of once per step,
//   * HBM traffic per particle-step drops by about K (the kernel stays far from the HBM roofline),
//   * slabs exchange their boundary bands once per K steps.
// The redundantly computed halo particles are bit-identical in every tile that computes them (IEEE arithmetic,
// canonical summation order), so "every tile keeps the particles that END inside it" partitions the particles
// exactly.
//
// Exactness of the halo argument.  Let d be a bound on the displacement of any particle in one step.  A tile
// loads every particle within H cells of its border (region R0).  Define U_0 = R0 and U_{s+1} = U_s shrunk by
// (cutoff + d).  Induction over the sub-steps: every loaded particle whose computed position at sub-step s lies in
// U_s carries its true state, and every particle truly in U_s is loaded (its force partners lie within one cutoff,
// i.e. inside U_s; a particle that shows up in U_{s+1} was within d of it one step earlier).  With
// H >= K * (1 + d / cutoff) the tile itself lies in U_K.  d is enforced, not assumed: every computed velocity
// component is compared with d / dt in every sub-step (all loaded particles, so that no mis-computed halo particle
// can jump into the trusted region either); a violation raises kErrSpeedBound, later launches become no-ops, and the
// host replays the batch from the untouched input buffers with K = 1 (d = H - 1 cells; kstep_after_sync) or, beyond that
// and for capacity overflows, hands the state over to the cellsort engine (psim_capi.cu: switch_to_cellsort).
// Ring culling: a halo particle r cells away from the tile matters only through sub-step H - 1 - r; the halo is sorted by
// ring on arrival and the passes of sub-step s stop at nproc[s].
//
// Layout in HBM.  Tiles of TS x TS cutoff cells; a tile owns a stripe of CAP slots in three streams pos (double2),
// vel (double2), id (int), double buffered by launch parity, plus acc (double2) for launches that keep the last
// accelerations.  Inside a stripe the particles are partitioned into nine classes by the H-wide border band their
// cell lies in, stored in ring order TL T TR R BR B BL L M, and a 16-int header per tile holds the ten prefix
// offsets.  A neighbour's halo band is then at most two contiguous slot ranges: a tile gathers its region with 10
// ranges (own stripe, four edge bands -- the E neighbour's left band wraps around the ring and takes two -- and four
// corners), each one TMA bulk copy (cp.async.bulk -> UBLKCP) for pos and one for vel, completion on an mbarrier.
//
// The kernel (kstep_kernel<TS, H, acc, peer>): persistent CTAs, one tile at a time, T threads.  Shared memory is what decides
// how many CTAs share an SM, and the other CTAs are what fills a tile's barrier waits and serial stretches, so the layout is
// lean (72 KB for a 64-cell tile: three CTAs of 256 threads per SM): ONE position array updated in place, velocities in load
// order, and one work area that is the cell table during the search, the force lists during the sub-steps and the landing zone
// of the next tile's positions during the store phase.
//   load     (by the last warp) the headers of the next tile are fetched while the current one computes.  Its POSITIONS are
//            bulk-copied into the work area the moment the current tile's last sub-step is over, so they fly during the store
//            phase (the loader warp is excused from it: named barrier of the other warps); its VELOCITIES go straight into the
//            velocity array as soon as the current tile has been stored and are first needed after the search.
//   arrive   positions leave the landing zone -- the halo ring-sorted (velocities stay in load order and are found through
//            horig) --, the cell table is wiped
//   bin      every loaded particle into a (TS+2H+4)^2 cell table (tile + halo + two empty guard rings) in shared memory: per cell
//            the head of a linked list (reference part3/gpu.cu:92-112 does this in global memory with 16 fixed slots) + one
//            occupancy bit
//   search   ONCE per tile: every particle walks the earlier half of its 5x5 cell neighbourhood (so each unordered pair is met
//            once) and lists the pairs within rs = cutoff + the distance two particles can close in the remaining sub-steps
//            (the speed bound is enforced below): the tile's candidate-pair list, ~1.25 pairs per particle.  Pairs that are
//            within the cutoff right now are handed to the first sub-step directly.  The cell table is not touched again.
//   K times  check   (sub-steps 2..K) the listed pairs are re-tested in exact FP64 against the cutoff (reference
//                    serial.cpp:24-26), one lane per pair; the in-range ones are collected tile-wide
//            eval    the in-range pairs are evaluated densely, one lane per pair, full warps: the sqrt + divisions of
//                    reference serial.cpp:29-33 run once per PAIR (the second particle takes the exact negative); the
//                    reference's 3x3 cell structure (serial.cpp:102-117) is re-imposed exactly here.  Each particle's
//                    force word counts its in-range neighbours and remembers its first two pairs
//            (slow)  only if some particle has >= 3 in-range neighbours: their sums in the canonical exact order
//            move    sum (<= 2 terms are order independent), move + reflect (serial.cpp:46-61) in place, speed check; after
//                    the last sub-step: final cell -> class
//   store    particles whose final cell lies in the tile are ranked inside their class (shared atomics), the class
//            offsets become the tile's new header, and pos / vel / id (/ acc) are stored to the other parity.  Tiles
//            of a slab's first / last row also store their facing band straight into the neighbour GPU's ghost row.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "psim_force.cuh"
#include "psim_internal.h"

namespace psim {

// ------------------------------------------------------------------------------------------
// compile-time tile configurations
// ------------------------------------------------------------------------------------------
// H halo cells, KMAX = H - 1 fused steps, T threads, CTAS resident CTAs per SM, CAP slots per stripe
// (mean population 0.2 TS^2), NMAX particles of one region (tile + halo) in shared memory.
template <int TS, int H> struct KCfg;
template <> struct KCfg<16, 3> { static constexpr int KMAX = 2, T = 64,  CTAS = 8, CAP = 128,  NMAX = 256; };
template <> struct KCfg<16, 4> { static constexpr int KMAX = 3, T = 64,  CTAS = 8, CAP = 128,  NMAX = 288; };
#ifndef PSIM_KSTEP_T32
#define PSIM_KSTEP_T32 128
#define PSIM_KSTEP_C32 5
#endif
template <> struct KCfg<32, 3> { static constexpr int KMAX = 2, T = PSIM_KSTEP_T32, CTAS = PSIM_KSTEP_C32, CAP = 352,  NMAX = 480; };
template <> struct KCfg<32, 4> { static constexpr int KMAX = 3, T = PSIM_KSTEP_T32, CTAS = PSIM_KSTEP_C32, CAP = 352,  NMAX = 512; };
// 64-cell tiles: three CTAs of 256 threads per SM (one CTA's shared memory is 72 KB).  A tile's phases are separated by
// barriers and contain serial stretches (the exact sqrt / division chain of the pair evaluation); what fills those gaps
// is the other CTAs of the SM -- measured: 1 CTA 17.2, 2 CTAs 31.2 G particle-steps/s at 256 threads (profiles/README.md)
#ifndef PSIM_KSTEP_T64
#define PSIM_KSTEP_T64 256
#define PSIM_KSTEP_C64 3
#endif
template <> struct KCfg<64, 3> { static constexpr int KMAX = 2, T = PSIM_KSTEP_T64, CTAS = PSIM_KSTEP_C64, CAP = 1024, NMAX = 1216; };
template <> struct KCfg<64, 4> { static constexpr int KMAX = 3, T = PSIM_KSTEP_T64, CTAS = PSIM_KSTEP_C64, CAP = 1024, NMAX = 1280; };
// one more tile size (halo 4 only), tunable at build time (tile-size sweeps: profiles/README.md)
#ifndef PSIM_KSTEP_TSX
#define PSIM_KSTEP_TSX 48
#define PSIM_KSTEP_TX 192
#define PSIM_KSTEP_CX 4
#define PSIM_KSTEP_CAPX 640
#define PSIM_KSTEP_NMAXX 800
#endif
constexpr int kTSX = PSIM_KSTEP_TSX;
template <> struct KCfg<kTSX, 4> { static constexpr int KMAX = 3, T = PSIM_KSTEP_TX, CTAS = PSIM_KSTEP_CX, CAP = PSIM_KSTEP_CAPX, NMAX = PSIM_KSTEP_NMAXX; };

constexpr int kHdrInts = 16;          // per tile: ten prefix offsets (ring order TL T TR R BR B BL L M, [9] = population)
constexpr int kRanges = 10;           // slot ranges that make up a tile's region
constexpr unsigned kNoOwner = 0xFFFFu;
constexpr int kGuard = 2;             // empty guard rings around the cell table: the candidate search looks two cells out
constexpr int kSlowMax = 16;          // in-range neighbours the canonical-order path sorts (beyond: hand-over to cellsort)
constexpr int kSlowSlots = 32;        // particles of one tile and sub-step that take the canonical-order path (beyond: hand-over)
// per-particle force word: count << 28 | second reference << 14 | first reference; a reference is the slot of an
// evaluated pair (13 bits) plus one bit that says "I am the pair's second particle: negate"
constexpr unsigned kRefBits = 14, kRefMask = (1u << kRefBits) - 1u, kRefNeg = 1u << 13, kCntShift = 28;

template <int TS, int H> struct KDims {
    using C = KCfg<TS, H>;
    static constexpr int TW = TS + 2 * H + 2 * kGuard;         // cell table side: tile + halo + guard rings (never occupied)
    static constexpr int NC = TW * TW, NC8 = (NC + 7) / 8 * 8;
    static constexpr int RW = (TW + 31) / 32 + 1;              // occupancy words per table row (+1: funnel shifts read one past)
    static constexpr int BM = (TW * RW + 3) / 4 * 4;           // the occupancy map, padded to 16 bytes
    static constexpr int NW = C::T / 32;
    static constexpr int PCAP = 2 * C::NMAX;                   // candidate pairs of one region (expected 1.25 per particle)
    static constexpr int NCON = C::NMAX / 2;                   // in-range pairs of one sub-step (expected 0.05 per particle)
    static constexpr int HREG = 2, HMAX = HREG * C::T;         // halo particles the ring sort handles (HREG per thread, in registers)
    static_assert(TS >= 2 * H && C::KMAX < H && C::NMAX < 0xFFF && C::T % 32 == 0 && C::CAP <= 0xFFF, "configuration");
    static_assert(TW <= 255, "cell codes pack row and column into 8 bits each");
    static_assert(NCON < (int)kRefNeg, "pair references are 13 bits");
    static constexpr int NLIST = NCON - kSlowSlots;            // ... of which the last kSlowSlots result slots serve the canonical-order path
    static_assert(PCAP >= 2 * NCON && NLIST > 32, "the first sub-step's in-range pairs grow down from the end of the candidate list");
};

template <int TS, int H> struct __align__(16) KSmem {
    using C = KCfg<TS, H>;
    using D = KDims<TS, H>;
    struct Tables {   // the cell table: everything the candidate search needs; dead once the search is over
        alignas(16) unsigned short head[D::NC8];   // per cell: first particle of its list (slot + 1, 0 = empty)
        alignas(16) unsigned bitmap[D::BM];        // one occupancy bit per cell
        alignas(16) unsigned short next[C::NMAX];  // list links (slot + 1, 0 = end)
    };
    struct Force {    // alive from the end of the search to the end of the last sub-step: shares the cell table's storage
        double2 wres[D::NCON];                     // contribution of evaluated pair g to its first particle (the second one takes the negative)
        unsigned inlist[D::NCON];                  // the in-range pairs of the current sub-step, waiting for their dense evaluation
    };
    struct Work {
        union alignas(16) {
            Tables t;
            Force f;
        } a;
        alignas(16) unsigned pairs[D::PCAP];       // candidate pairs i | j << 16 (from the front); the pairs already in range at the first
    };                                             // sub-step are listed a second time from the back (the search cannot use `inlist` yet)
    double2 pos[C::NMAX];                    // positions, updated in place (a sub-step's phases are separated by barriers)
    double2 vel[C::NMAX];                    // velocities in LOAD order (TMA destination): a ring-sorted halo particle finds its own through horig
    union alignas(16) {
        Work w;
        double2 pland[C::NMAX];              // landing zone of the NEXT tile's positions while this tile is stored (TMA destination)
    } u;
    unsigned pw[C::NMAX];                    // bin -> search: table row << 8 | column;  sub-steps: force word;  store phase: class << 12 | rank
    unsigned short wslow[D::NW][kSlowMax];   // per warp: neighbours of a particle on the canonical-order path (rank << 12 | slot)
    unsigned short horig[D::HMAX];           // ring-sorted halo particle -> its slot in load order (velocity, id look-up at store time)
    unsigned long long mbar_p, mbar_v;       // TMA completion: positions of the next tile, velocities of this one
    int rsrc[2][kRanges], rlen[2][kRanges], rdst[2][kRanges];   // slot ranges of this tile / the next: global slot, length, first shared slot
    int ncount[2];                           // particles of this tile's / the next tile's region
    int nproc[8];                            // per sub-step: particles that are still processed (own + rings that still matter)
    int ringcnt[8];
    int segcnt[12], segoff[12];
    int npairs, ncon, work;
    int nslow[2], nslowslot[2];              // by sub-step parity: particles with >= 3 in-range neighbours, result slots handed out
    int flags, hw_region, hw_stripe, hw_pairs;
};

static_assert(KCfg<64, 4>::CTAS * (sizeof(KSmem<64, 4>) + 1024) <= 233472 && KCfg<64, 3>::CTAS * (sizeof(KSmem<64, 3>) + 1024) <= 233472,
              "the CTAs of the 64-cell kernel must fit one SM's shared memory");
static_assert(sizeof(KSmem<64, 4>::Force) <= sizeof(KSmem<64, 4>::Tables), "the force lists of a 64-cell tile fit inside its cell table");
static_assert(KCfg<kTSX, 4>::CTAS * (sizeof(KSmem<kTSX, 4>) + 1024) <= 233472, "the CTAs of the extra tile size must fit one SM's shared memory");

// the ten ranges: neighbour (dr, dc) and the classes [cb, ce) of ITS stripe that lie within H cells of this tile
__constant__ signed char kRangeTab[kRanges][4] = {
    {0, 0, 0, 9},     // own stripe
    {-1, 0, 4, 7},    // N neighbour: its bottom band  BR B BL
    {1, 0, 0, 3},     // S neighbour: its top band     TL T TR
    {0, -1, 2, 5},    // W neighbour: its right band   TR R BR
    {0, 1, 6, 8},     // E neighbour: its left band    BL L ...
    {0, 1, 0, 1},     //                               ... TL (the ring wraps)
    {-1, -1, 4, 5},   // NW: BR corner
    {-1, 1, 6, 7},    // NE: BL corner
    {1, -1, 2, 3},    // SW: TR corner
    {1, 1, 0, 1},     // SE: TL corner
};

struct KParams {
    const double2 *pos_in, *vel_in;
    const int *id_in, *hdr_in;
    double2 *pos_out, *vel_out, *acc_out;
    int *id_out, *hdr_out;
    double2* acc_tmp;     // per CTA scratch (NMAX entries): accelerations of the last sub-step until the slots are known
    int ntx, nty;         // tiles per side (global)
    int tr_base;          // global tile row of local row 0
    int lrow0, row_stride, ntiles;   // tile t of the launch lies in local row lrow0 + (t / ntx) * row_stride
    int bincnt;
    double size;
    int nsub;             // steps fused in this launch (0: only re-partition the stripes)
    double vlim2;         // square of the speed bound that keeps the halo argument and the candidate list valid for nsub steps
    double rs2;           // square of the candidate radius: cutoff + twice the distance the speed bound allows in nsub - 1 steps
    int* err;
    int seq;              // launch number (reported with the first error)
    int ringsort;         // order the halo by ring and stop processing rings that no longer matter
    // slabs with peer-memory exchange: the first / last owned tile row also stores its facing band and headers into the
    // neighbour GPU's ghost row (pointers to slot 0 / header 0 of that ROW, parity written)
    double2 *peer_pos[2], *peer_vel[2];
    int *peer_id[2], *peer_hdr[2];
    int last_lrow;
};

// ---- PTX helpers: mbarrier + 1-D bulk copy (TMA) --------------------------------------------------
__device__ __forceinline__ unsigned k_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void k_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(k_smem_u32(bar)), "r"(parity), "r"(200000u)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void k_tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(k_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(k_smem_u32(bar))
                 : "memory");
}
// one hardware shared-memory atomic per lane (the compiler would otherwise wrap atomicAdd into a ballot / leader / shuffle
// sequence of a dozen instructions; the few lanes of a diverged warp that get here are cheaper served by the LSU)
__device__ __forceinline__ int k_atoms_add(int* p, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(k_smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void k_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// the K-step kernel
// ------------------------------------------------------------------------------------------
// ---- per-particle helpers --------------------------------------------------------------------------------------------
// A particle's table cell: its exact cell, clamped to the box (reference semantics for x == size) and to the table interior
// (only particles whose computed state is already worthless can leave the loaded region).
template <int TW>
__device__ __forceinline__ void k_table_cell(double x, double y, int rbase, int cbase, bool at_wall, int bincnt, int& lr, int& lc) {
    int gr = __double2int_rd(div_by_bin(x)), gc = __double2int_rd(div_by_bin(y));
    if (at_wall) {
        gr = min(max(gr, 0), bincnt - 1);
        gc = min(max(gc, 0), bincnt - 1);
    }
    lr = min(max(gr - rbase, kGuard), TW - 1 - kGuard);
    lc = min(max(gc - cbase, kGuard), TW - 1 - kGuard);
}
// insert particle p into the list of its cell (two 16-bit heads share a word: exchange one half, keep the other)
template <int TS, int H>
__device__ __forceinline__ void k_bin_particle(KSmem<TS, H>& S, int p, int lr, int lc) {
    constexpr int TW = KDims<TS, H>::TW, RW = KDims<TS, H>::RW;
    const int cell = lr * TW + lc;
    unsigned* w = reinterpret_cast<unsigned*>(S.u.w.a.t.head) + (cell >> 1);
    const unsigned sh = (unsigned)(cell & 1) << 4;
    const unsigned ent = (unsigned)(p + 1) << sh, keep = ~(0xFFFFu << sh);
    unsigned old = *w;
    for (;;) {
        const unsigned prev = atomicCAS(w, old, (old & keep) | ent);
        if (prev == old) break;
        old = prev;
    }
    S.u.w.a.t.next[p] = (unsigned short)((old >> sh) & 0xFFFFu);
    S.pw[p] = (unsigned)((lr << 8) | lc);
    atomicOr(&S.u.w.a.t.bitmap[lr * RW + (lc >> 5)], 1u << (lc & 31));
}
// final cell -> owner test, class, rank inside the class
template <int TS, int H>
__device__ __forceinline__ void k_classify(KSmem<TS, H>& S, int p, int lr, int lc) {
    const int er = lr - (H + kGuard), ec = lc - (H + kGuard);
    unsigned oc = kNoOwner;
    const bool mine = (unsigned)er < (unsigned)TS && (unsigned)ec < (unsigned)TS;
    const int rb = er < H ? 0 : (er >= TS - H ? 2 : 1), cb = ec < H ? 0 : (ec >= TS - H ? 2 : 1);
    const unsigned cls = (unsigned)((0x456387210ull >> (4 * (rb * 3 + cb))) & 0xFull);   // ring order TL T TR R BR B BL L M
    // most particles are interior (class M): one shared atomic per warp for them instead of one per lane on the same word
    const unsigned act = __activemask();
    const unsigned mm = __ballot_sync(act, mine && cls == 8u);
    if (mine) {
        int rank;
        if (cls == 8u) {
            const int leader = __ffs((int)mm) - 1, lane = (int)(threadIdx.x & 31u);
            int base = 0;
            if (lane == leader) base = atomicAdd(&S.segcnt[8], __popc(mm));
            rank = __shfl_sync(mm, base, leader) + __popc(mm & ((1u << lane) - 1u));
        } else {
            rank = atomicAdd(&S.segcnt[cls], 1);
        }
        oc = (cls << 12) | (unsigned)min(rank, 0xFFF);
    }
    S.pw[p] = oc;
}

// Walls (reference serial.cpp:53-61), out of line and through shared memory: only tiles whose region touches a wall get here,
// and keeping the bounce loops out of the move phase keeps their registers out of it too.
// (Out-of-line functions name the shared-memory block themselves instead of taking pointers into it: a generic pointer to
// shared memory makes the compiler rebuild the CTA's shared window address -- S2R SR_CgaCtaId -- all over the hot loops.)
template <int TS, int H>
static __device__ __noinline__ void k_reflect_in_place(int p, int vi, double size) {
    extern __shared__ __align__(128) unsigned char k_smem_raw[];
    KSmem<TS, H>& S = *reinterpret_cast<KSmem<TS, H>*>(k_smem_raw);
    double2 q = S.pos[p], v = S.vel[vi];
    reflect_particle(q.x, q.y, v.x, v.y, size);
    S.pos[p] = q;
    S.vel[vi] = v;
}

// Are the cells of two particles the same or adjacent (the reference only ever compares a particle with the members of its
// 3x3 cell neighbourhood, serial.cpp:102-117)?  Two particles within the cutoff practically always are; the test only decides
// pairs that sit within a few ulps of two cell edges at once.  Returns the neighbour's visit rank as seen from `a`, or -1.
__device__ __forceinline__ int k_neighbour_rank(const double2 a, const double2 c, int bincnt) {
    const int dr = axis_cell(c.x, bincnt) - axis_cell(a.x, bincnt), dc = axis_cell(c.y, bincnt) - axis_cell(a.y, bincnt);
    if (dr < -1 || dr > 1 || dc < -1 || dc > 1) return -1;
    return visit_rank(dr, dc);
}

// ---- candidate search (once per tile) ---------------------------------------------------------------------------------------
// Every loaded particle looks at the "earlier half" of its 5x5 cell neighbourhood -- the two rows above, the two cells to its
// left, and the members of its own cell that follow it in the cell's list -- so that every unordered pair is met exactly once,
// and lists the pairs within rs (cutoff + twice the distance a particle may travel in the remaining sub-steps of this launch,
// which the speed check enforces): no other pair can come within the cutoff before the next launch.  The reference's 3x3
// walk (serial.cpp:102-117) is the rs = cutoff special case; its cell structure is re-imposed exactly when a pair is evaluated.
// Pairs that are within the cutoff right now are listed a second time, from the END of the candidate array downwards: the first
// sub-step evaluates them without a check pass (`inlist` shares its storage with the cell table the search is still reading).
template <int TS, int H>
static __device__ __forceinline__ void k_search(KSmem<TS, H>& S, int n, int np0, double rs2) {
    using C = KCfg<TS, H>;
    using D = KDims<TS, H>;
    constexpr int TW = D::TW, RW = D::RW;
    const int lane = threadIdx.x & 31;
    const double2* posb = S.pos;
    const unsigned short* next = S.u.w.a.t.next;
    const int nch = (n + 31) >> 5;
    // chunks of 32 particles are handed out dynamically (their cost varies with the local density); the warp reconverges
    // after every chunk -- without the explicit __syncwarp the lanes drift apart and every later instruction is issued
    // several times for a few lanes each
#pragma unroll 1
    for (;;) {
        int ch = 0;
        if (lane == 0) ch = k_atoms_add(&S.work, 1);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= nch) break;
        const int p = ch * 32 + lane;
        if (p < n) {
        const unsigned cc = S.pw[p];
        S.pw[p] = 0u;   // from here on: the particle's force word
        const int lr = (int)(cc >> 8), lc = (int)(cc & 0xFFu);
        const double2 me = posb[p];
        // occupancy of rows lr-2, lr-1 (five cells from column lc-2) and of row lr (columns lc-2, lc-1, and my own cell)
        const unsigned* rb = S.u.w.a.t.bitmap + (lr - 2) * RW + ((lc - 2) >> 5);
        const unsigned sh = (unsigned)(lc - 2) & 31u;
        unsigned m = (__funnelshift_r(rb[0], rb[1], sh) & 31u) | ((__funnelshift_r(rb[RW], rb[RW + 1], sh) & 31u) << 8) |
                     ((__funnelshift_r(rb[2 * RW], rb[2 * RW + 1], sh) & 3u) << 16);
        const unsigned own_next = next[p];
        if (own_next) m |= 1u << 18;   // members of my own cell that follow me in its list
        const unsigned short* hcorner = S.u.w.a.t.head + (lr - 2) * TW + (lc - 2);
#pragma unroll 1
        while (m) {
            const int k = __ffs((int)m) - 1;
            m &= m - 1;
            unsigned h = k == 18 ? own_next : (unsigned)hcorner[(k >> 3) * TW + (k & 7)];
#pragma unroll 1
            do {   // the occupancy bit guarantees a non-empty list
                const unsigned j = h - 1u;
                const double2 pj = posb[j];
                h = next[j];
                const double dx = __dsub_rn(pj.x, me.x), dy = __dsub_rn(pj.y, me.y);
                const double r2 = pair_r2(dx, dy);
                if (r2 <= rs2 && min((unsigned)p, j) < (unsigned)np0) {
                    const unsigned ij = (unsigned)p | (j << 16);
                    const int slot = k_atoms_add(&S.npairs, 1);
                    if (slot < D::PCAP) S.u.w.pairs[slot] = ij;
                    if (!(r2 > kCutoff2) && r2 != 0.0) {   // in range right now: the first sub-step evaluates it
                        const int g = k_atoms_add(&S.ncon, 1);
                        if (g < D::NLIST) S.u.w.pairs[D::PCAP - 1 - g] = ij;
                    }
                }
            } while (h);
        }
        }
        __syncwarp();
    }
}

// ---- a sub-step, first half: which listed pairs are within the cutoff now; their contributions ---------------------------------
// k_pair_check collects the in-range pairs of the whole tile in one list (the first sub-step gets them from the search itself);
// after a barrier k_pair_eval evaluates them densely, one lane per pair, full warps: the sqrt + divisions of reference
// serial.cpp:29-33 run once per PAIR (the second particle takes the exact negative).  Each particle's force word counts its
// in-range neighbours and remembers the first two evaluated pairs.
template <int TS, int H>
static __device__ __forceinline__ void k_pair_check(KSmem<TS, H>& S, int np, int nvalid) {
    using C = KCfg<TS, H>;
    using D = KDims<TS, H>;
    constexpr int T = C::T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int npairs = min(S.npairs, D::PCAP);
#pragma unroll 1
    for (int base = warp * 32; base < npairs; base += T) {
        const int e = base + lane;
        bool hit = false;
        unsigned ij = 0;
        if (e < npairs) {
            ij = S.u.w.pairs[e];
            const unsigned i = ij & 0xFFFFu, j = ij >> 16;
            const double2 a = S.pos[i], c = S.pos[j];
            const double dx = __dsub_rn(c.x, a.x), dy = __dsub_rn(c.y, a.y);
            const double r2 = pair_r2(dx, dy);
            // a pair matters while one of the two is still processed and both still carry valid positions; pairs at
            // distance exactly 0 contribute coef * 0 = -0, which a sum from +0 absorbs (serial.cpp:107 meets the self pair)
            hit = !(r2 > kCutoff2) && r2 != 0.0 && min(i, j) < (unsigned)np && max(i, j) < (unsigned)nvalid;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
            int gbase = 0;
            if (lane == 0) gbase = atomicAdd(&S.ncon, __popc(bal));
            gbase = __shfl_sync(0xffffffffu, gbase, 0);
            const int g = gbase + __popc(bal & lt_mask);
            if (hit && g < D::NLIST) S.u.w.a.f.inlist[g] = ij;
        }
    }
}

// `list` / `dir`: the in-range pairs -- inlist upwards, or (first sub-step) the tail of the candidate array downwards
template <int TS, int H>
static __device__ __forceinline__ void k_pair_eval(KSmem<TS, H>& S, const unsigned* list, int dir, int np, int bincnt, int sidx) {
    using C = KCfg<TS, H>;
    using D = KDims<TS, H>;
    const double2* posb = S.pos;
    const int ncon = min(S.ncon, D::NLIST);
#pragma unroll 1
    for (int g = threadIdx.x; g < ncon; g += C::T) {
        const unsigned ij = list[dir * g];
        const unsigned i = ij & 0xFFFFu, j = ij >> 16;
        const double2 a = posb[i], c = posb[j];
        const double dx = __dsub_rn(c.x, a.x), dy = __dsub_rn(c.y, a.y);
        // two particles less than 0.9999 cells apart along both axes lie in the same or in adjacent cells (the computed cell
        // index is off by far less than 1e-4 cells); the exact test only runs for the others
        bool adjacent = true;
        if (fmax(fabs(dx), fabs(dy)) > 0.9999 * kBin) adjacent = k_neighbour_rank(a, c, bincnt) >= 0;
        if (adjacent) {
            double cx, cy;
            pair_contrib(dx, dy, pair_r2(dx, dy), cx, cy);
            S.u.w.a.f.wres[g] = make_double2(cx, cy);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const unsigned q = side ? j : i;
                if (q < (unsigned)np) {   // (a partner that is no longer processed only acts on the other one)
                    const unsigned cnt = atomicAdd(&S.pw[q], 1u << kCntShift) >> kCntShift;
                    const unsigned ref = (unsigned)g | (side ? kRefNeg : 0u);
                    if (cnt < 2u) atomicOr(&S.pw[q], ref << (cnt * kRefBits));
                    if (cnt == 2u) atomicAdd(&S.nslow[sidx], 1);            // a third neighbour: the canonical-order pass must run
                    if (cnt >= 14u) atomicOr(&S.flags, kErrSmemOverflow);   // the 4-bit count would wrap: hand over
                }
            }
        }
    }
}

// Rare path: a particle with three or more in-range neighbours.  The whole warp scans the pair list for its partners; the owner
// lane then sums their contributions in ascending (reference cell-visit rank, x, y) order -- the order the oracle uses
// (oracle/psim_oracle.c, "Summation order").
static __device__ __forceinline__ double2 kslow_sum(const double2* xy, unsigned short* nb, int n, int i) {
    auto less = [&](unsigned u, unsigned v) {
        if ((u >> 12) != (v >> 12)) return (u >> 12) < (v >> 12);
        const double2 pu = xy[u & 0xFFFu], pv = xy[v & 0xFFFu];
        if (pu.x != pv.x) return pu.x < pv.x;
        return pu.y < pv.y;
    };
    for (int a = 1; a < n; ++a) {
        const unsigned cur = nb[a];
        int q = a;
        while (q > 0 && less(cur, nb[q - 1])) {
            nb[q] = nb[q - 1];
            --q;
        }
        nb[q] = (unsigned short)cur;
    }
    const double2 me = xy[i];
    double sx = 0.0, sy = 0.0;
    for (int a = 0; a < n; ++a) {
        const double2 pj = xy[nb[a] & 0xFFFu];
        const double dx = __dsub_rn(pj.x, me.x), dy = __dsub_rn(pj.y, me.y);
        double cx, cy;
        pair_contrib(dx, dy, pair_r2(dx, dy), cx, cy);
        sx = __dadd_rn(sx, cx);
        sy = __dadd_rn(sy, cy);
    }
    return make_double2(sx, sy);
}

template <int TS, int H>
static __device__ __noinline__ double2 kslow_force(unsigned slow, int p, int nvalid, int bincnt) {
    using D = KDims<TS, H>;
    extern __shared__ __align__(128) unsigned char k_smem_raw[];
    KSmem<TS, H>& S = *reinterpret_cast<KSmem<TS, H>*>(k_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const double2* xy = S.pos;
    const int npairs = min(S.npairs, D::PCAP);   // (the first sub-step's list at the tail never overlaps: checked after the search)
    double2 out = make_double2(0.0, 0.0);
    while (slow) {
        const int leader = __ffs((int)slow) - 1;
        slow &= slow - 1;
        const unsigned i = (unsigned)__shfl_sync(0xffffffffu, p, leader);
        const double2 me = xy[i];
        int cnt = 0;
        for (int base = 0; base < npairs; base += 32) {
            const int e = base + lane;
            bool hit = false;
            unsigned ent = 0;
            if (e < npairs) {
                const unsigned ij = S.u.w.pairs[e];
                const unsigned a = ij & 0xFFFFu, c = ij >> 16;
                if ((a == i || c == i) && max(a, c) < (unsigned)nvalid) {
                    const unsigned o = a == i ? c : a;
                    const double2 po = xy[o];
                    const double dx = __dsub_rn(po.x, me.x), dy = __dsub_rn(po.y, me.y);
                    const double r2 = pair_r2(dx, dy);
                    const int rank = k_neighbour_rank(me, po, bincnt);
                    hit = !(r2 > kCutoff2) && r2 != 0.0 && rank >= 0;
                    ent = ((unsigned)max(rank, 0) << 12) | o;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int k = cnt + __popc(bal & lt_mask);
                if (k < kSlowMax) S.wslow[warp][k] = (unsigned short)ent;
            }
            cnt += __popc(bal);
        }
        __syncwarp();
        if (lane == leader) {
            if (cnt > kSlowMax) atomicOr(&S.flags, kErrSmemOverflow);   // denser than any physical configuration: hand over
            out = kslow_sum(xy, S.wslow[warp], min(cnt, kSlowMax), (int)i);
        }
        __syncwarp();
    }
    return out;
}

// ---- canonical-order pass (only in sub-steps where some particle has three or more in-range neighbours) -------------------------
// Runs between the pair evaluation and the move phase, i.e. while every position is still the sub-step's input (the move phase
// updates positions in place).  The sum lands in one of the last kSlowSlots result slots and the particle's force word is rewritten
// to "one neighbour, that slot": +0 + sum == sum exactly (a sum that starts from +0 is never -0).
template <int TS, int H>
static __device__ __noinline__ void k_slow_phase(int np, int nvalid, int bincnt, int sidx) {
    using D = KDims<TS, H>;
    extern __shared__ __align__(128) unsigned char k_smem_raw[];
    KSmem<TS, H>& S = *reinterpret_cast<KSmem<TS, H>*>(k_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nch = (np + 31) >> 5;
    for (int ch = warp; ch < nch; ch += D::NW) {
        const int p = ch * 32 + lane;
        const unsigned cnt = (p < np ? S.pw[p] : 0u) >> kCntShift;
        const unsigned slow = __ballot_sync(0xffffffffu, cnt >= 3u);
        if (!slow) continue;
        const double2 a = kslow_force<TS, H>(slow, p, nvalid, bincnt);
        if (cnt >= 3u) {
            const int k = atomicAdd(&S.nslowslot[sidx], 1);
            if (k < kSlowSlots) {
                const int slot = D::NCON - 1 - k;
                S.u.w.a.f.wres[slot] = a;
                S.pw[p] = (1u << kCntShift) | (unsigned)slot;
            } else {
                atomicOr(&S.flags, kErrSmemOverflow);   // hand over
                S.pw[p] = 0u;
            }
        }
    }
}

// ---- a sub-step, second half: sum, move, speed check; after the last one: final cell -> class ---------------------------------
// Positions are updated in place (only the owner lane of a particle touches them here; the check / evaluation phases that read
// other particles' positions are separated from this one by barriers).  A ring-sorted halo particle finds its velocity, which
// stays in load order, through horig.
template <int TS, int H, bool kStoreAcc>
static __device__ __forceinline__ void k_move_phase(KSmem<TS, H>& S, int np, int n_own, bool sorted, bool last, int rbase, int cbase,
                                                    bool at_wall, int bincnt, double size, double vlim2, double2* acc_tmp) {
    using C = KCfg<TS, H>;
    using D = KDims<TS, H>;
    constexpr int TW = D::TW, NW = D::NW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nch = (np + 31) >> 5;
    bool too_fast = false;
#pragma unroll 1
    for (int ch = warp; ch < nch; ch += NW) {
        const int p = ch * 32 + lane;
        const unsigned w = p < np ? S.pw[p] : 0u;
        const unsigned cnt = w >> kCntShift;
        double ax = 0.0, ay = 0.0;
        if (p < np) {
            if (cnt != 0u) {   // 1 or 2 (the canonical-order pass has turned every larger count into "1, its result slot")
                // (the pair's second particle takes the exact negative: flip the sign bit)
                auto signed_of = [](double v, unsigned ref) {
                    return __hiloint2double(__double2hiint(v) ^ (int)((ref & kRefNeg) << 18), __double2loint(v));
                };
                const double2 c0 = S.u.w.a.f.wres[w & (kRefNeg - 1u)];
                ax = __dadd_rn(ax, signed_of(c0.x, w));
                ay = __dadd_rn(ay, signed_of(c0.y, w));
                if (cnt == 2u) {
                    const unsigned w1 = w >> kRefBits;
                    const double2 c1 = S.u.w.a.f.wres[w1 & (kRefNeg - 1u)];
                    ax = __dadd_rn(ax, signed_of(c1.x, w1));
                    ay = __dadd_rn(ay, signed_of(c1.y, w1));
                }
            }
            // integrate (reference serial.cpp:46-51); walls (serial.cpp:53-61) only where the loaded region touches one:
            // elsewhere a particle that reaches a wall has left the trusted part of the region anyway
            const int vi = (sorted && p >= n_own) ? (int)S.horig[p - n_own] : p;
            const double2 me = S.pos[p];
            double2 v = S.vel[vi];
            v.x = __dadd_rn(v.x, __dmul_rn(ax, kDt));
            v.y = __dadd_rn(v.y, __dmul_rn(ay, kDt));
            double x = __dadd_rn(me.x, __dmul_rn(v.x, kDt)), y = __dadd_rn(me.y, __dmul_rn(v.y, kDt));
            too_fast |= !(__fma_rn(v.x, v.x, __dmul_rn(v.y, v.y)) < vlim2);
            S.pos[p] = make_double2(x, y);
            S.vel[vi] = v;
            if (at_wall) {
                k_reflect_in_place<TS, H>(p, vi, size);
                const double2 q = S.pos[p];
                x = q.x;
                y = q.y;
            }
            if (!last) {
                S.pw[p] = 0u;
            } else {
                int lr, lc;
                k_table_cell<TW>(x, y, rbase, cbase, at_wall, bincnt, lr, lc);
                k_classify<TS, H>(S, p, lr, lc);
                if (kStoreAcc) acc_tmp[p] = make_double2(ax, ay);
            }
        }
    }
    if (__any_sync(0xffffffffu, too_fast) && lane == 0) atomicOr(&S.flags, kErrSpeedBound);
}


template <int TS, int H, bool kStoreAcc, bool kPeer>
__global__ void __launch_bounds__(KCfg<TS, H>::T, KCfg<TS, H>::CTAS) kstep_kernel(const KParams P) {
    using C = KCfg<TS, H>;
    using D = KDims<TS, H>;
    constexpr int T = C::T, CAP = C::CAP, NMAX = C::NMAX, TW = D::TW, RW = D::RW, NW = D::NW;

    extern __shared__ __align__(128) unsigned char k_smem_raw[];
    KSmem<TS, H>& S = *reinterpret_cast<KSmem<TS, H>*>(k_smem_raw);

    // an earlier launch hit a capacity / speed bound: leave the state untouched so that the host can replay it
    if (*reinterpret_cast<volatile int*>(P.err + 8) != 0) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsub = P.nsub;
    const int G = gridDim.x;

    if (tid == 0) {
        k_mbar_init(&S.mbar_p, 1);
        k_mbar_init(&S.mbar_v, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.flags = 0;
        S.hw_region = S.hw_stripe = S.hw_pairs = 0;
    }
    if (tid < 9) S.segcnt[tid] = 0;
    if (tid < 8) S.ringcnt[tid] = 0;
    // ---- loader (the last warp) --------------------------------------------------------------------------------------------
    // Slot range `lane` (< kRanges) of the region of launch tile t: length and first global slot.  The headers of a tile are
    // fetched one tile before its bulk copies are issued; its positions are copied while the previous tile is stored, its
    // velocities as soon as the previous tile's have been stored.
    // (the loader is the LAST warp: it has the fewest particle chunks; tiles are walked without divisions)
    constexpr int kLoader = NW - 1;
    const int walk_q = (G / P.ntx) * P.row_stride, walk_r = G % P.ntx;
    struct Walk {
        int t, lrow, tc;
    };
    auto advance = [&](Walk& w) {
        w.t += G;
        w.tc += walk_r;
        w.lrow += walk_q;
        if (w.tc >= P.ntx) {
            w.tc -= P.ntx;
            w.lrow += P.row_stride;
        }
    };
    Walk cur{(int)blockIdx.x, P.lrow0 + ((int)blockIdx.x / P.ntx) * P.row_stride, (int)blockIdx.x % P.ntx};
    Walk fw = cur;   // the tile whose headers are fetched next
    auto fetch_range = [&](int& len, int& src) {
        len = 0;
        src = 0;
        if (fw.t < P.ntiles && lane < kRanges && (lane == 0 || nsub > 0)) {
            const int rt_dr = kRangeTab[lane][0], rt_dc = kRangeTab[lane][1], rt_cb = kRangeTab[lane][2], rt_ce = kRangeTab[lane][3];
            const int ntr = P.tr_base + fw.lrow + rt_dr, ntc = fw.tc + rt_dc;
            if (ntr >= 0 && ntr < P.nty && ntc >= 0 && ntc < P.ntx) {
                const int nlt = (fw.lrow + rt_dr) * P.ntx + ntc;
                const int* h = P.hdr_in + (size_t)nlt * kHdrInts;
                len = max(min(h[rt_ce], CAP) - min(h[rt_cb], CAP), 0);
                src = nlt * CAP + min(h[rt_cb], CAP);
            }
        }
        advance(fw);
    };
    // Publish the ranges of a tile in slot set q and bulk-copy its positions into the landing zone (the storage of the cell
    // table / force lists / candidate pairs, dead between the last sub-step of a tile and the wipe of the next one).
    auto issue_positions = [&](int len, int src, int q) {
        int inc = len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += a;
        }
        int dst = inc - len;
        const int total = __shfl_sync(0xffffffffu, inc, kRanges - 1);
        if (total > NMAX) {   // region does not fit: truncate (memory safety) and report
            dst = min(dst, NMAX);
            len = min(len, NMAX - dst);
        }
        const int n = min(total, NMAX);
        if (lane < kRanges) {
            S.rsrc[q][lane] = src;
            S.rlen[q][lane] = len;
            S.rdst[q][lane] = dst;
        }
        if (lane == 0) {
            S.ncount[q] = n;
            if (total > NMAX) atomicOr(&S.flags, kErrSmemOverflow);
            S.hw_region = max(S.hw_region, total);
            k_mbar_arrive_expect_tx(&S.mbar_p, (unsigned)n * 16u);
        }
        __syncwarp();
        if (len > 0) k_tma_load_1d(&S.u.pland[dst], P.pos_in + src, (unsigned)len * 16u, &S.mbar_p);
    };
    // The velocities of the tile whose ranges are published in slot set q go straight into S.vel (load order).
    auto issue_velocities = [&](int q) {
        int len = 0, src = 0, dst = 0;
        if (lane < kRanges) {
            len = S.rlen[q][lane];
            src = S.rsrc[q][lane];
            dst = S.rdst[q][lane];
        }
        if (lane == 0) k_mbar_arrive_expect_tx(&S.mbar_v, (unsigned)S.ncount[q] * 16u);
        __syncwarp();
        if (len > 0) k_tma_load_1d(&S.vel[dst], P.vel_in + src, (unsigned)len * 16u, &S.mbar_v);
    };
    int next_len = 0, next_src = 0;
    __syncthreads();   // mbarriers initialised
    if (warp == kLoader) {
        fetch_range(next_len, next_src);
        issue_positions(next_len, next_src, 0);
        fetch_range(next_len, next_src);
    }

    for (int it = 0; cur.t < P.ntiles; ++it, advance(cur)) {
        const int q = it & 1;
        const int lrow = cur.lrow, tc = cur.tc;
        const int rbase = (P.tr_base + lrow) * TS - H - kGuard, cbase = tc * TS - H - kGuard;   // global cell of table row / column 0
        const bool at_wall = rbase + kGuard <= 0 || cbase + kGuard <= 0 || rbase + TW - 1 - kGuard >= P.bincnt - 1 ||
                             cbase + TW - 1 - kGuard >= P.bincnt - 1;

        // a particle's table cell: its exact cell, clamped to the box (reference semantics for x == size) and to the table
        // interior (only particles whose computed state is already worthless can leave the loaded region)
        auto table_cell = [&](double x, double y, int& lr, int& lc) { k_table_cell<TW>(x, y, rbase, cbase, at_wall, P.bincnt, lr, lc); };

        // the previous tile's velocities have been stored (barrier at the end of the loop body): this tile's may land
        if (warp == kLoader) issue_velocities(q);

        // ---- arrival: positions leave the landing zone, the halo is ordered by ring ------------------------------------
        k_mbar_wait(&S.mbar_p, (unsigned)q);
        const int n = S.ncount[q];
        const int n_own = S.rlen[q][0], n_halo = n - n_own;
        for (int p = tid; p < n_own; p += T) S.pos[p] = S.u.pland[p];
        // A halo particle at ring r (cells between it and the tile, 1..H) can influence the tile's final state only through
        // sub-step H - 1 - r, and nothing reads it after sub-step H - r: sorted by ring, the particles that still matter are a
        // prefix of the region, and the passes of sub-step s simply stop at n_proc[s].  (Skipped when the halo exceeds HREG
        // particles per thread.)  Ring and class counters were zeroed at the end of the previous tile.  Only positions are
        // moved: the velocities stay in load order and are found through horig.
        const bool ringsort = P.ringsort && nsub > 0 && n_halo <= D::HMAX;
        double2 hpos[D::HREG];
        int hcode[D::HREG];
        if (ringsort) {
#pragma unroll
            for (int k = 0; k < D::HREG; ++k) {
                const int h = tid + k * T;
                hcode[k] = 0;
                hpos[k] = make_double2(0.0, 0.0);
                if (h < n_halo) {
                    hpos[k] = S.u.pland[n_own + h];
                    int lr, lc;
                    table_cell(hpos[k].x, hpos[k].y, lr, lc);
                    const int er = lr - (H + kGuard), ec = lc - (H + kGuard);
                    const int dr = er < 0 ? -er : (er >= TS ? er - TS + 1 : 0), dc = ec < 0 ? -ec : (ec >= TS ? ec - TS + 1 : 0);
                    const int ring = min(max(max(dr, dc), 1), H);
                    hcode[k] = (ring << 16) | atomicAdd(&S.ringcnt[ring], 1);
                }
            }
        } else {
            for (int p = n_own + tid; p < n; p += T) S.pos[p] = S.u.pland[p];
        }
        __syncthreads();   // the landing zone is free (it becomes the cell table again); ring counts complete
        if (ringsort) {
#pragma unroll
            for (int k = 0; k < D::HREG; ++k) {
                const int h = tid + k * T;
                if (h < n_halo) {
                    const int ring = hcode[k] >> 16;
                    int o = hcode[k] & 0xFFFF;
                    for (int r = 1; r < ring; ++r) o += S.ringcnt[r];
                    S.pos[n_own + o] = hpos[k];
                    S.horig[o] = (unsigned short)(n_own + h);
                }
            }
            if (tid < C::KMAX) {   // sub-step tid processes rings <= H - 1 - tid
                int c = n_own;
                for (int r = 1; r <= H - 1 - tid; ++r) c += S.ringcnt[r];
                S.nproc[tid] = c;
            }
        } else if (tid < C::KMAX) {
            S.nproc[tid] = n;
        }
        {   // wipe the cell table and the occupancy map
            uint4* hq = reinterpret_cast<uint4*>(S.u.w.a.t.head);
            for (int c = tid; c < D::NC8 / 8; c += T) hq[c] = make_uint4(0u, 0u, 0u, 0u);
            uint4* bq = reinterpret_cast<uint4*>(S.u.w.a.t.bitmap);
            for (int c = tid; c < D::BM / 4; c += T) bq[c] = make_uint4(0u, 0u, 0u, 0u);
            if (tid == 0) S.npairs = S.ncon = S.work = S.nslow[0] = S.nslowslot[0] = 0;
        }
        __syncthreads();

        if (nsub == 0) {
            for (int p = tid; p < n; p += T) {
                const double2 a = S.pos[p];
                int lr, lc;
                table_cell(a.x, a.y, lr, lc);
                k_classify<TS, H>(S, p, lr, lc);
            }
        } else {
            for (int p = tid; p < n; p += T) {
                const double2 a = S.pos[p];
                int lr, lc;
                table_cell(a.x, a.y, lr, lc);
                k_bin_particle<TS, H>(S, p, lr, lc);
            }
        }
        k_fence_proxy_async();   // (only matters for nsub == 0: the barrier below is then the last one before the next bulk copies)
        __syncthreads();

        // ---- candidate pairs of the whole launch, then the fused time steps ---------------------------------------------
        if (nsub > 0) {
            k_search<TS, H>(S, n, S.nproc[0], P.rs2);
            __syncthreads();
            // the candidate list (from the front) and the first sub-step's in-range list (from the back) must not meet
            if (tid == 0 && S.npairs + min(S.ncon, D::NCON) > D::PCAP) atomicOr(&S.flags, kErrSmemOverflow);
        }
        k_mbar_wait(&S.mbar_v, (unsigned)q);   // the velocities have had the whole arrival / binning / search to land
        for (int s = 0; s < nsub; ++s) {
            const int np = S.nproc[s], nvalid = s == 0 ? n : S.nproc[s - 1];
            if (s > 0) {   // (the first sub-step's in-range pairs come from the search)
                k_pair_check<TS, H>(S, np, nvalid);
                __syncthreads();
                k_pair_eval<TS, H>(S, S.u.w.a.f.inlist, 1, np, P.bincnt, s & 1);
            } else {
                k_pair_eval<TS, H>(S, S.u.w.pairs + (D::PCAP - 1), -1, np, P.bincnt, 0);
            }
            __syncthreads();
            if (tid == 0) {
                if (S.ncon > D::NLIST) atomicOr(&S.flags, kErrSmemOverflow);
                S.hw_pairs = max(S.hw_pairs, S.ncon);
                S.ncon = 0;
                S.nslow[(s + 1) & 1] = S.nslowslot[(s + 1) & 1] = 0;   // (the counters of this sub-step's parity are still being read)
            }
            if (S.nslow[s & 1] != 0) {   // same value in every thread: written before the barrier above, reset one sub-step later
                k_slow_phase<TS, H>(np, nvalid, P.bincnt, s & 1);
                __syncthreads();
            }
            k_move_phase<TS, H, kStoreAcc>(S, np, n_own, ringsort, s + 1 == nsub, rbase, cbase, at_wall, P.bincnt, P.size, P.vlim2,
                                           P.acc_tmp + (size_t)blockIdx.x * NMAX);
            if (s + 1 == nsub) k_fence_proxy_async();   // my accesses to the tables / lists precede the next tile's bulk copies into their storage
            __syncthreads();
        }

        // ---- store phase: the loader warp issues the next tile's position copies while the other warps store this tile -------
        if (warp == kLoader) {
            if (cur.t + G < P.ntiles) {
                issue_positions(next_len, next_src, q ^ 1);
                fetch_range(next_len, next_src);
            }
        } else {
            constexpr int TS_ = T - 32;   // storing threads (named barrier 1)
            // the ids of the tile's own stripe (shared slots [0, n_own)) are requested now, all at once: their latency hides
            // behind the header work and the barrier below instead of being paid once per loop iteration
            constexpr int kIdRegs = (CAP + TS_ - 1) / TS_;
            const int own_src = S.rsrc[q][0];
            int idreg[kIdRegs];
#pragma unroll
            for (int k = 0; k < kIdRegs; ++k) idreg[k] = tid + k * TS_ < n_own ? P.id_in[own_src + tid + k * TS_] : 0;
            // class offsets -> header, particles -> the other parity's stripe
            if (warp == 0) {
                int v = lane < 9 ? S.segcnt[lane] : 0, inc = v;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const int a = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += a;
                }
                if (lane < 10) S.segoff[lane] = inc - v;   // exclusive prefix; [9] = population
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TS_) : "memory");
            const int lt = lrow * P.ntx + tc;
            const size_t gbase = (size_t)lt * CAP;
            double2* ppos = nullptr;
            double2* pvel = nullptr;
            int *pid = nullptr, *phdr = nullptr;
            unsigned peer_classes = 0;   // bit c: class c also goes to the neighbour slab
            if (kPeer) {
                // first owned row -> lower neighbour (needs my top band TL T TR); last owned row -> upper neighbour (BR B BL).
                // A one-row slab would have to serve both: such slabs use the NCCL flavour (kstep host code).
                const int side = lrow == 1 ? 0 : (lrow == P.last_lrow ? 1 : -1);
                if (side >= 0 && P.peer_pos[side]) {
                    ppos = P.peer_pos[side] + (size_t)tc * CAP;
                    pvel = P.peer_vel[side] + (size_t)tc * CAP;
                    pid = P.peer_id[side] + (size_t)tc * CAP;
                    phdr = P.peer_hdr[side] + (size_t)tc * kHdrInts;
                    peer_classes = side == 0 ? 0x007u : 0x070u;
                }
            }
            if (tid < 10) {
                const int o = min(S.segoff[tid], CAP);
                P.hdr_out[(size_t)lt * kHdrInts + tid] = o;
                if (kPeer && phdr) phdr[tid] = o;
                if (tid == 9) {
                    if (S.segoff[9] > CAP) atomicOr(&S.flags, kErrTileOverflow);
                    S.hw_stripe = max(S.hw_stripe, S.segoff[9]);
                }
            }
            const int nfin = nsub > 0 ? S.nproc[nsub - 1] : n;   // particles beyond were never candidates for the tile
            auto store_one = [&](int p, int own_id) {
                const unsigned oc = S.pw[p];
                if (oc == kNoOwner) return;
                const unsigned cls = oc >> 12;
                const int d = S.segoff[cls] + (int)(oc & 0xFFFu);
                if (d >= CAP) return;
                int id = own_id, vi = p;
                if (p >= n_own) {   // a halo particle that moved into the tile: find its source slot
                    if (ringsort) vi = (int)S.horig[p - n_own];   // its slot in load order
                    id = 0;
                    for (int g = 1; g < kRanges; ++g) {
                        const int o = vi - S.rdst[q][g];
                        if ((unsigned)o < (unsigned)S.rlen[q][g]) id = P.id_in[S.rsrc[q][g] + o];
                    }
                }
                const double2 pq = S.pos[p], v = S.vel[vi];
                P.pos_out[gbase + d] = pq;
                P.vel_out[gbase + d] = v;
                P.id_out[gbase + d] = id;
                if (kStoreAcc) P.acc_out[gbase + d] = nsub > 0 ? P.acc_tmp[(size_t)blockIdx.x * NMAX + p] : make_double2(0.0, 0.0);
                if (kPeer && ((peer_classes >> cls) & 1u)) {
                    ppos[d] = pq;
                    pvel[d] = v;
                    pid[d] = id;
                }
            };
#pragma unroll
            for (int k = 0; k < kIdRegs; ++k)
                if (tid + k * TS_ < nfin) store_one(tid + k * TS_, idreg[k]);
#pragma unroll 1
            for (int p = tid + kIdRegs * TS_; p < nfin; p += TS_) store_one(p, 0);   // (beyond the stripe capacity: halo particles only)
            k_fence_proxy_async();   // my reads of the velocities precede the next tile's bulk copy into S.vel
            asm volatile("bar.sync 1, %0;" ::"n"(TS_) : "memory");   // class counters have been read by everybody
            if (tid < 9) S.segcnt[tid] = 0;
            if (tid < 8) S.ringcnt[tid] = 0;
        }
        __syncthreads();   // the tile is completely stored: positions, velocities and codes may be overwritten
    }
    // ---- report ------------------------------------------------------------------------------------------------------
    __syncthreads();
    if (tid == 0) {
        if (S.flags) {
            atomicOr(P.err, S.flags);
            atomicCAS(P.err + 8, 0, P.seq + 1);   // later launches become no-ops: the input of launch `seq` stays intact
        }
        atomicMax(P.err + 3, S.hw_stripe);
        atomicMax(P.err + 4, S.hw_region);
        atomicMax(P.err + 7, S.hw_pairs);
    }
}

// ------------------------------------------------------------------------------------------
// initial tiling and observation
// ------------------------------------------------------------------------------------------
// one thread per input record: owned tile rows only; arrival order inside a stripe is arbitrary (forces are summed in a
// canonical order); the stripes are partitioned into classes by a 0-step launch afterwards
__global__ void __launch_bounds__(256) kstep_fill_kernel(const particle_t* __restrict__ p, int n, int id0, int bincnt, int ts,
                                                         int cap, int ntx, int tr_begin, int tr_end, int tr_base,
                                                         double2* __restrict__ pos, double2* __restrict__ vel,
                                                         int* __restrict__ sid, int* __restrict__ tcount, int* __restrict__ err) {
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

// ---- cooperative upload (slabs): every rank uploads 1/nranks of the caller's array and the records travel to their slab over
// NVLink, instead of every rank uploading everything and throwing 1 - 1/nranks of it away ----------------------------------------
// destination slab of every record of my chunk (+ per-destination counts); tr_bounds[r] = first tile row of rank r, [nranks] = ntx
__global__ void __launch_bounds__(256) kstep_route_count_kernel(const particle_t* __restrict__ p, int m, int bincnt, int ts, int nranks,
                                                                const int* __restrict__ tr_bounds, int* __restrict__ dest,
                                                                int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int tr = axis_cell(p[i].x, bincnt) / ts;
    int r = 0;
    while (r + 1 < nranks && tr >= tr_bounds[r + 1]) ++r;
    dest[i] = r;
    atomicAdd(counts + r, 1);
}
// records grouped by destination: rec = {x y vx vy}, ids = original index
__global__ void __launch_bounds__(256) kstep_route_pack_kernel(const particle_t* __restrict__ p, int m, int id0, const int* __restrict__ dest,
                                                               const int* __restrict__ offs, int* __restrict__ cursor,
                                                               double4* __restrict__ rec, int* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int r = dest[i];
    const int d = offs[r] + atomicAdd(cursor + r, 1);
    const double2* q = reinterpret_cast<const double2*>(p + i);
    const double2 a = q[0], b = q[1];
    rec[d] = make_double4(a.x, a.y, b.x, b.y);
    ids[d] = id0 + i;
}
// received records -> tile stripes (the counterpart of kstep_fill_kernel)
__global__ void __launch_bounds__(256) kstep_fill_records_kernel(const double4* __restrict__ rec, const int* __restrict__ ids, int n, int bincnt,
                                                                 int ts, int cap, int ntx, int tr_begin, int tr_end, int tr_base,
                                                                 double2* __restrict__ pos, double2* __restrict__ vel, int* __restrict__ sid,
                                                                 int* __restrict__ tcount, int* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 q = rec[i];
    const int tr = axis_cell(q.x, bincnt) / ts, tc = axis_cell(q.y, bincnt) / ts;
    if (tr < tr_begin || tr >= tr_end) {   // routed to the wrong slab: cannot happen, but never write outside my rows
        atomicOr(err, kErrLostParticle);
        return;
    }
    const int lt = (tr - tr_base) * ntx + tc;
    const int slot = atomicAdd(tcount + lt, 1);
    if (slot >= cap) {
        atomicOr(err, kErrTileOverflow);
        return;
    }
    const size_t d = (size_t)lt * cap + slot;
    pos[d] = make_double2(q.x, q.y);
    vel[d] = make_double2(q.z, q.w);
    sid[d] = ids[i];
}

// headers of the freshly filled (unpartitioned) stripes: everything counts as class M
__global__ void __launch_bounds__(256) kstep_hdr_init_kernel(const int* __restrict__ tcount, int ntiles, int cap, int* __restrict__ hdr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntiles * kHdrInts) return;
    const int c = i % kHdrInts;
    hdr[i] = c == 9 ? min(tcount[i / kHdrInts], cap) : 0;
}

// gather the owned particles into a compact SoA (observation calls); one CTA per owned tile
__global__ void __launch_bounds__(128) kstep_gather_kernel(const double2* __restrict__ pos, const double2* __restrict__ vel,
                                                           const int* __restrict__ sid, const int* __restrict__ hdr,
                                                           const double2* __restrict__ acc, int cap, int ntx, bool have_acc,
                                                           double* __restrict__ gx, double* __restrict__ gy,
                                                           double* __restrict__ gvx, double* __restrict__ gvy,
                                                           double* __restrict__ gax, double* __restrict__ gay,
                                                           int* __restrict__ gid, int* __restrict__ cursor) {
    __shared__ int s_base;
    const int lt = (1 + blockIdx.x / ntx) * ntx + blockIdx.x % ntx;   // owned rows start at local row 1
    const int n = min(hdr[(size_t)lt * kHdrInts + 9], cap);
    if (threadIdx.x == 0) s_base = n ? atomicAdd(cursor, n) : 0;
    __syncthreads();
    const int base = s_base;
    const size_t gbase = (size_t)lt * cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 p = pos[gbase + i], v = vel[gbase + i];
        const double2 a = have_acc ? acc[gbase + i] : make_double2(0.0, 0.0);
        gx[base + i] = p.x;
        gy[base + i] = p.y;
        gvx[base + i] = v.x;
        gvy[base + i] = v.y;
        gax[base + i] = a.x;
        gay[base + i] = a.y;
        gid[base + i] = sid[gbase + i];
    }
}

// Write-back in ORIGINAL particle order straight from the stripes (the drivers' view of `parts`, reference
// part1/main.cpp:135-136 / part3/main.cu:134-136): out[id] = {x y vx vy ax ay}, or only xy[id] = {x y}.
__global__ void __launch_bounds__(128) kstep_writeback_kernel(const double2* __restrict__ pos, const double2* __restrict__ vel,
                                                              const int* __restrict__ sid, const int* __restrict__ hdr,
                                                              const double2* __restrict__ acc, int cap, int ntx, bool have_acc,
                                                              particle_t* __restrict__ out, double2* __restrict__ out_xy) {
    const int lt = (1 + blockIdx.x / ntx) * ntx + blockIdx.x % ntx;
    const int n = min(hdr[(size_t)lt * kHdrInts + 9], cap);
    const size_t gbase = (size_t)lt * cap;
    if (out_xy) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) out_xy[sid[gbase + i]] = pos[gbase + i];
    } else {
        // three adjacent lanes write the three 16-byte pieces of one 48-byte record: one contiguous transaction per record
        // (this matters when `out` is pinned host memory written over PCIe)
        for (int u = threadIdx.x; u < 3 * n; u += blockDim.x) {
            const int i = u / 3, piece = u - 3 * i;
            const int id = sid[gbase + i];
            const double2 v = piece == 0 ? pos[gbase + i] : piece == 1 ? vel[gbase + i] : (have_acc ? acc[gbase + i] : make_double2(0.0, 0.0));
            reinterpret_cast<double2*>(out + id)[piece] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct KLaunchRec {   // what the host needs to replay a launch
    int seq, nsub, parity_in;
    bool store;
    long long steps_before;
};

struct KstepEngine {
    DeviceArena mem;
    int ts = 0, h = 0, kmax = 0, cap = 0, nmax = 0, threads = 0, ctas_per_sm = 1;
    size_t smem = 0;
    int sms = 148;
    int ntx = 0;
    int tr_begin = 0, tr_end = 0, lrows = 0, lrows_alloc = 0;
    int ksteps = 4;                 // steps fused per launch (PSIM_KSTEPS, <= kmax)
    int ringsort = 1;               // PSIM_RINGSORT=0 disables the halo ring culling
    // stripes + headers, double buffered by launch parity; one allocation per parity (its CUDA IPC handle is shared with
    // the neighbour slabs): [headers | pos | vel | id]
    char* buf[2] = {nullptr, nullptr};
    size_t buf_bytes = 0, off_pos = 0, off_vel = 0, off_id = 0;
    double2 *pos[2] = {nullptr, nullptr}, *vel[2] = {nullptr, nullptr}, *acc = nullptr, *acc_tmp = nullptr;
    int *sid[2] = {nullptr, nullptr}, *hdr[2] = {nullptr, nullptr};
    int* tcount = nullptr;
    int parity = 0;
    bool acc_valid = false;
    bool ghost_fresh = false;
    int grid_cap = 0;
    int seq = 0;                    // launches issued so far
    std::vector<KLaunchRec> log;    // launches since the last successful synchronisation
    int recoveries = 0;
    // cooperative upload: the fill is deferred to psim_comm_connect (the caller's host array must stay valid until then)
    const particle_t* pending_parts = nullptr;
    int pending_n = 0;
    // gather scratch
    DeviceArena gmem;
    SoAView g{};
    int* g_cursor = nullptr;
    int g_capacity = 0;
};

void tiled_slab_rows(int ntx, int rank, int nranks, int* begin, int* end);   // psim_tiled.cu
int tiled_tile_rows(int bincnt, int ts);

static KParams kmake_params(psim_sim* sim, KstepEngine* e, int parity_in) {
    KParams P{};
    const int po = parity_in ^ 1;
    P.pos_in = e->pos[parity_in]; P.vel_in = e->vel[parity_in]; P.id_in = e->sid[parity_in]; P.hdr_in = e->hdr[parity_in];
    P.pos_out = e->pos[po]; P.vel_out = e->vel[po]; P.id_out = e->sid[po]; P.hdr_out = e->hdr[po];
    P.acc_out = e->acc;
    P.acc_tmp = e->acc_tmp;
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
    P.ringsort = e->ringsort;
    for (int side = 0; side < 2; ++side) {
        P.peer_pos[side] = P.peer_vel[side] = nullptr;
        P.peer_id[side] = P.peer_hdr[side] = nullptr;
    }
    if (sim->p2p) {
        // the neighbour's buffer has the same [headers | pos | vel | id] layout; only its row count differs
        for (int side = 0; side < 2; ++side) {
            const int nb = side == 0 ? sim->rank - 1 : sim->rank + 1;
            if (nb < 0 || nb >= sim->nranks || !sim->peer_exports[side][po]) continue;
            int b, en;
            tiled_slab_rows(e->ntx, nb, sim->nranks, &b, &en);
            const size_t nb_rows_alloc = (size_t)(en - b) + 2;
            const size_t grow = side == 0 ? (size_t)(en - b) + 1 : 0;   // its upper ghost row / its lower ghost row
            char* base = sim->peer_exports[side][po];
            const size_t nb_tiles = nb_rows_alloc * e->ntx, nb_slots = nb_tiles * e->cap;
            auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
            const size_t o_pos = up(nb_tiles * kHdrInts * sizeof(int)), o_vel = o_pos + up(nb_slots * sizeof(double2)),
                         o_id = o_vel + up(nb_slots * sizeof(double2));
            P.peer_hdr[side] = reinterpret_cast<int*>(base) + grow * e->ntx * kHdrInts;
            P.peer_pos[side] = reinterpret_cast<double2*>(base + o_pos) + grow * e->ntx * e->cap;
            P.peer_vel[side] = reinterpret_cast<double2*>(base + o_vel) + grow * e->ntx * e->cap;
            P.peer_id[side] = reinterpret_cast<int*>(base + o_id) + grow * e->ntx * e->cap;
        }
    }
    return P;
}

template <int TS, int H>
static int klaunch(psim_sim* sim, KstepEngine* e, KParams& P, bool store_acc, cudaStream_t s) {
    const int grid = std::min(P.ntiles, e->grid_cap);
    const bool peer = P.peer_pos[0] || P.peer_pos[1];
    constexpr int kThreads = KCfg<TS, H>::T;
    constexpr size_t kSmem = sizeof(KSmem<TS, H>);
    if (store_acc) {
        if (peer) kstep_kernel<TS, H, true, true><<<grid, kThreads, kSmem, s>>>(P);
        else kstep_kernel<TS, H, true, false><<<grid, kThreads, kSmem, s>>>(P);
    } else {
        if (peer) kstep_kernel<TS, H, false, true><<<grid, kThreads, kSmem, s>>>(P);
        else kstep_kernel<TS, H, false, false><<<grid, kThreads, kSmem, s>>>(P);
    }
    ++sim->launches;
    return PSIM_OK;
}

// one launch over local tile rows lrow0, lrow0 + stride, ... (nrows of them)
static int klaunch_rows(psim_sim* sim, KstepEngine* e, int parity_in, int nsub, bool store_acc, int seq, int lrow0, int nrows,
                        int row_stride, bool allow_peer, cudaStream_t s) {
    if (nrows <= 0) return PSIM_OK;
    KParams P = kmake_params(sim, e, parity_in);
    if (!allow_peer)
        for (int side = 0; side < 2; ++side) P.peer_pos[side] = nullptr;
    P.lrow0 = lrow0;
    P.row_stride = row_stride;
    P.ntiles = nrows * e->ntx;
    if (allow_peer && sim->nranks > 1) P.acc_tmp += (size_t)e->grid_cap * e->nmax;   // the boundary launch's own scratch
    P.nsub = nsub;
    P.seq = seq;
    // Speed bound of the launch (2 % margins for the rounding of the region arithmetic).  Halo: H cells allow nsub fused steps
    // a displacement of d = (H - nsub) / nsub cells per step.  Candidate list: it is built from the positions at the first
    // sub-step with radius rs < 2 cells (the search covers the 5x5 cell neighbourhood) and must contain every pair that comes
    // within the cutoff in the nsub - 1 steps that follow: rs = cutoff + 2 (nsub - 1) v dt.
    {
        double v = 1e300, rs = PSIM_CUTOFF;
        if (nsub > 0) v = 0.98 * ((double)(e->h - nsub) / nsub) * PSIM_BIN_SIZE / PSIM_DT;
        if (nsub > 1) {
            v = std::min(v, 0.98 * PSIM_BIN_SIZE / (2.0 * (nsub - 1) * PSIM_DT));
            rs = PSIM_CUTOFF + 2.0 * (nsub - 1) * v * PSIM_DT * (1.0 + 1e-9);
        }
        P.vlim2 = v * v;
        P.rs2 = rs * rs * (1.0 + 1e-12);
    }
    switch (e->ts * 8 + e->h) {
        case 16 * 8 + 3: return klaunch<16, 3>(sim, e, P, store_acc, s);
        case 16 * 8 + 4: return klaunch<16, 4>(sim, e, P, store_acc, s);
        case 32 * 8 + 3: return klaunch<32, 3>(sim, e, P, store_acc, s);
        case 32 * 8 + 4: return klaunch<32, 4>(sim, e, P, store_acc, s);
        case 64 * 8 + 3: return klaunch<64, 3>(sim, e, P, store_acc, s);
        case 64 * 8 + 4: return klaunch<64, 4>(sim, e, P, store_acc, s);
        case kTSX * 8 + 4: return klaunch<kTSX, 4>(sim, e, P, store_acc, s);
    }
    return fail(PSIM_ERR_INVALID, "kstep tile size %d / halo %d not instantiated", e->ts, e->h);
}

template <int TS, int H>
static int kconfigure(KstepEngine* e) {
    e->h = H;
    e->kmax = KCfg<TS, H>::KMAX;
    e->cap = KCfg<TS, H>::CAP;
    e->nmax = KCfg<TS, H>::NMAX;
    e->threads = KCfg<TS, H>::T;
    e->smem = sizeof(KSmem<TS, H>);
    auto prepare = [&](auto kernel) -> int {
        PSIM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
        PSIM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        return PSIM_OK;
    };
    PSIM_TRY(prepare(kstep_kernel<TS, H, true, true>));
    PSIM_TRY(prepare(kstep_kernel<TS, H, true, false>));
    PSIM_TRY(prepare(kstep_kernel<TS, H, false, true>));
    PSIM_TRY(prepare(kstep_kernel<TS, H, false, false>));
    int per_sm = 0;
    PSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kstep_kernel<TS, H, false, false>, KCfg<TS, H>::T, e->smem));
    e->ctas_per_sm = std::max(1, per_sm);
    if (const char* cap = std::getenv("PSIM_CTAS_PER_SM")) {   // tuning / profiling knob
        const int c = std::atoi(cap);
        if (c >= 1) e->ctas_per_sm = std::min(e->ctas_per_sm, c);
    }
    return PSIM_OK;
}

bool kstep_tile_supported(int ts) { return ts == 16 || ts == 32 || ts == 64 || ts == kTSX; }

int kstep_default_tile(int bincnt) {
    // 64-cell tiles amortise the halo best (region / tile = 1.34); smaller boxes take smaller tiles so that there are
    // enough tiles to spread over the SMs
    const long long n64 = (bincnt + 63) / 64, n32 = (bincnt + 31) / 32;
    if (n64 * n64 >= 592) return 64;
    if (n32 * n32 >= 296) return 32;
    return bincnt >= 64 ? 32 : 16;
}

int kstep_finish_fill(psim_sim* sim, bool* unsuitable);

int kstep_create(psim_sim* sim, const psim_config* cfg, const particle_t* parts, int n, bool parts_on_device, bool* unsuitable) {
    *unsuitable = false;
    cudaStream_t s = sim->stream;
    static const bool trace = std::getenv("PSIM_TRACE") != nullptr;   // phase wall times of create on stderr
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        const double t = now();
        std::fprintf(stderr, "[psim trace] kstep_create rank %d: %-28s %8.3f ms\n", sim->rank, what, 1e3 * (t - t_prev));
        t_prev = t;
    };
    int ts = cfg->tile_cells;
    if (ts == 0) ts = kstep_default_tile(sim->bincnt);
    if (!kstep_tile_supported(ts)) return fail(PSIM_ERR_INVALID, "tile_cells must be 16, 32, %d or 64 (got %d)", kTSX, ts);
    auto* e = new KstepEngine();
    sim->kstep = e;
    e->ts = ts;
    PSIM_CUDA(cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, sim->device));
    int halo = 4;   // 4 halo cells / up to 3 fused steps; PSIM_HALO=3 selects 3 cells / up to 2 steps
    if (const char* hv = std::getenv("PSIM_HALO")) halo = std::atoi(hv);
    if (halo != 3 && halo != 4) return fail(PSIM_ERR_INVALID, "PSIM_HALO must be 3 or 4 (got %d)", halo);
    if (ts == 16) PSIM_TRY(halo == 3 ? (kconfigure<16, 3>(e)) : (kconfigure<16, 4>(e)));
    if (ts == 32) PSIM_TRY(halo == 3 ? (kconfigure<32, 3>(e)) : (kconfigure<32, 4>(e)));
    if (ts == 64) PSIM_TRY(halo == 3 ? (kconfigure<64, 3>(e)) : (kconfigure<64, 4>(e)));
    if (ts == kTSX) {
        if (halo != 4) return fail(PSIM_ERR_INVALID, "%d-cell tiles come with a 4-cell halo only", kTSX);
        PSIM_TRY((kconfigure<kTSX, 4>(e)));
    }
    e->grid_cap = e->sms * e->ctas_per_sm;
    e->ksteps = e->kmax;
    if (const char* r = std::getenv("PSIM_RINGSORT")) e->ringsort = std::atoi(r) != 0;
    if (const char* k = std::getenv("PSIM_KSTEPS")) e->ksteps = std::min(std::max(std::atoi(k), 1), e->kmax);
    if (cfg->steps_per_launch > 0) e->ksteps = std::min(cfg->steps_per_launch, e->kmax);
    e->ntx = tiled_tile_rows(sim->bincnt, ts);
    if (sim->nranks > e->ntx) return fail(PSIM_ERR_INVALID, "more slabs (%d) than tile rows (%d)", sim->nranks, e->ntx);
    tiled_slab_rows(e->ntx, sim->rank, sim->nranks, &e->tr_begin, &e->tr_end);
    e->lrows = e->tr_end - e->tr_begin;
    e->lrows_alloc = e->lrows + 2;
    sim->row_begin = e->tr_begin * ts;
    sim->row_end = std::min(e->tr_end * ts, sim->bincnt);

    const size_t tiles = (size_t)e->lrows_alloc * e->ntx, slots = tiles * e->cap;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    e->off_pos = up(tiles * kHdrInts * sizeof(int));
    e->off_vel = e->off_pos + up(slots * sizeof(double2));
    e->off_id = e->off_vel + up(slots * sizeof(double2));
    e->buf_bytes = e->off_id + up(slots * sizeof(int));
    for (int b = 0; b < 2; ++b) {
        PSIM_TRY(e->mem.alloc(&e->buf[b], e->buf_bytes));
        e->hdr[b] = reinterpret_cast<int*>(e->buf[b]);
        e->pos[b] = reinterpret_cast<double2*>(e->buf[b] + e->off_pos);
        e->vel[b] = reinterpret_cast<double2*>(e->buf[b] + e->off_vel);
        e->sid[b] = reinterpret_cast<int*>(e->buf[b] + e->off_id);
        PSIM_CUDA(cudaMemsetAsync(e->hdr[b], 0, tiles * kHdrInts * sizeof(int), s));
    }
    PSIM_TRY(e->mem.alloc(&e->acc, slots));
    PSIM_TRY(e->mem.alloc(&e->acc_tmp, (size_t)2 * e->grid_cap * e->nmax));   // (x2: a slab's boundary and interior launches run concurrently)
    PSIM_TRY(e->mem.alloc(&e->tcount, tiles));
    PSIM_CUDA(cudaMemsetAsync(e->tcount, 0, sizeof(int) * tiles, s));
    if (trace)
        std::fprintf(stderr, "[psim trace] kstep_create rank %d: %d-cell tiles, halo %d, %d threads x %d CTAs per SM, %zu bytes of shared memory per CTA\n",
                     sim->rank, e->ts, e->h, e->threads, e->ctas_per_sm, e->smem);
    lap("configure + allocate");
    // Slabs with a host array: by default the upload is cooperative and happens in psim_comm_connect (kstep_distributed_fill)
    static const bool coop = !(std::getenv("PSIM_COOP_UPLOAD") && std::getenv("PSIM_COOP_UPLOAD")[0] == '0');
    if (sim->nranks > 1 && !parts_on_device && coop && n > 0) {
        e->pending_parts = parts;
        e->pending_n = n;
        e->parity = 0;
        return PSIM_OK;
    }
    // fill: device input is read in place; host input is streamed through two bounded staging buffers on two auxiliary
    // streams, so that the scatter of chunk k runs while chunk k + 1 is on the wire (the copies of both streams share the one
    // host->device engine; stream order protects each staging buffer)
    if (parts_on_device) {
        if (n > 0) {
            kstep_fill_kernel<<<(n + 255) / 256, 256, 0, s>>>(parts, n, 0, sim->bincnt, ts, e->cap, e->ntx, e->tr_begin, e->tr_end,
                                                              e->tr_begin - 1, e->pos[0], e->vel[0], e->sid[0], e->tcount, sim->d_err);
            ++sim->launches;
        }
        PSIM_CUDA(cudaGetLastError());
    } else if (n > 0) {
        DeviceArena stage;
        particle_t* d_stage[2] = {nullptr, nullptr};
        cudaStream_t aux[2] = {nullptr, nullptr};
        cudaEvent_t ev = nullptr;
        const int chunk = std::min(n, 2 << 20);
        const int nbuf = n > chunk ? 2 : 1;
        int st = PSIM_OK;
        cudaError_t ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        for (int b = 0; b < nbuf && st == PSIM_OK && ce == cudaSuccess; ++b) {
            st = stage.alloc(&d_stage[b], (size_t)chunk);
            if (st == PSIM_OK) ce = cudaStreamCreateWithFlags(&aux[b], cudaStreamNonBlocking);
        }
        if (st == PSIM_OK && ce == cudaSuccess) ce = cudaEventRecord(ev, s);   // the zeroed tile counters
        for (int b = 0; b < nbuf && st == PSIM_OK && ce == cudaSuccess; ++b) ce = cudaStreamWaitEvent(aux[b], ev, 0);
        int k = 0;
        for (int off = 0; off < n && st == PSIM_OK && ce == cudaSuccess; off += chunk, ++k) {
            const int m = std::min(chunk, n - off), b = k % nbuf;
            ce = cudaMemcpyAsync(d_stage[b], parts + off, sizeof(particle_t) * (size_t)m, cudaMemcpyHostToDevice, aux[b]);
            if (ce != cudaSuccess) break;
            kstep_fill_kernel<<<(m + 255) / 256, 256, 0, aux[b]>>>(d_stage[b], m, off, sim->bincnt, ts, e->cap, e->ntx, e->tr_begin, e->tr_end,
                                                                   e->tr_begin - 1, e->pos[0], e->vel[0], e->sid[0], e->tcount, sim->d_err);
            ++sim->launches;
            ce = cudaGetLastError();
        }
        for (int b = 0; b < nbuf; ++b) {   // everything below is ordered behind both auxiliary streams; the staging is freed after them
            if (!aux[b]) continue;
            if (ce == cudaSuccess) ce = cudaEventRecord(ev, aux[b]);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s, ev, 0);
            const cudaError_t ce2 = cudaStreamSynchronize(aux[b]);
            if (ce == cudaSuccess) ce = ce2;
            cudaStreamDestroy(aux[b]);
        }
        if (ev) cudaEventDestroy(ev);
        stage.release();
        PSIM_TRY(st);
        if (ce != cudaSuccess) return fail(PSIM_ERR_CUDA, "kstep_create: upload: %s", cudaGetErrorString(ce));
    }
    lap("upload + fill");
    return kstep_finish_fill(sim, unsuitable);
}

// after the stripes of parity 0 are filled: capacity check, headers, partition launch
int kstep_finish_fill(psim_sim* sim, bool* unsuitable) {
    KstepEngine* e = sim->kstep;
    cudaStream_t s = sim->stream;
    const int ts = e->ts;
    const size_t tiles = (size_t)e->lrows_alloc * e->ntx;
    // suitability: the densest tile must leave headroom for fluctuations, else the caller falls back
    {
        std::vector<int> h(tiles);
        PSIM_CUDA(cudaMemcpyAsync(h.data(), e->tcount, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, s));
        PSIM_CUDA(cudaStreamSynchronize(s));
        int worst = 0;
        for (int v : h) worst = std::max(worst, v);
        // the region (tile + halo) of the densest neighbourhood must fit shared memory too: bound it by the area ratio
        const double region_ratio = (double)(ts + 2 * e->h) * (ts + 2 * e->h) / ((double)ts * ts);
        if (worst + worst / 8 + 16 > e->cap || worst * region_ratio * 1.05 + 16 > e->nmax) {
            *unsuitable = true;
            PSIM_CUDA(cudaMemsetAsync(sim->d_err, 0, sizeof(int), s));
            return fail(PSIM_ERR_UNSUPPORTED, "kstep engine: densest %dx%d-cell tile holds %d particles (stripe capacity %d, region capacity %d)",
                        ts, ts, worst, e->cap, e->nmax);
        }
    }
    kstep_hdr_init_kernel<<<((int)tiles * kHdrInts + 255) / 256, 256, 0, s>>>(e->tcount, (int)tiles, e->cap, e->hdr[0]);
    ++sim->launches;
    // partition the stripes into classes: a 0-step launch from parity 0 into parity 1
    PSIM_TRY(klaunch_rows(sim, e, 0, 0, true, e->seq++, 1, e->lrows, 1, false, s));
    PSIM_CUDA(cudaGetLastError());
    e->parity = 1;
    e->acc_valid = true;   // zeros
    return PSIM_OK;
}

int kstep_exchange(psim_sim* sim, int parity, cudaStream_t s);   // psim_comm.cpp

bool kstep_fill_pending(psim_sim* sim) { return sim->kstep && sim->kstep->pending_parts != nullptr; }

// Cooperative upload, called from psim_comm_connect once the communicator exists: rank r uploads records
// [r n / nranks, (r+1) n / nranks) of the caller's array (every rank was handed the same array, like the reference's
// MPI_Bcast, part2/main.cpp:150), routes them by slab on the device and sends every slab its share over NCCL.
int kstep_distributed_fill(psim_sim* sim) {
    KstepEngine* e = sim->kstep;
    cudaStream_t s = sim->stream;
    const int n = e->pending_n, R = sim->nranks;
    const particle_t* parts = e->pending_parts;
    e->pending_parts = nullptr;
    const int lo = (int)((long long)n * sim->rank / R), hi = (int)((long long)n * (sim->rank + 1) / R), m = hi - lo;
    static const bool trace = std::getenv("PSIM_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        const double t = now();
        std::fprintf(stderr, "[psim trace] distributed fill rank %d: %-24s %8.3f ms\n", sim->rank, what, 1e3 * (t - t_prev));
        t_prev = t;
    };
    DeviceArena tmp;
    particle_t* d_chunk = nullptr;
    double4 *d_rec = nullptr, *d_recv = nullptr;
    int *d_dest = nullptr, *d_ids = nullptr, *d_recv_ids = nullptr, *d_counts = nullptr, *d_bounds = nullptr;
    std::vector<int> bounds((size_t)R + 1), matrix((size_t)R * R, 0), offs((size_t)R + 1, 0);
    for (int r = 0; r < R; ++r) {
        int b, en;
        tiled_slab_rows(e->ntx, r, R, &b, &en);
        bounds[(size_t)r] = b;
        bounds[(size_t)R] = en;
    }
    int st = PSIM_OK;
    auto cuda_ok = [&](cudaError_t err, const char* what) {
        if (err != cudaSuccess && st == PSIM_OK) st = fail(PSIM_ERR_CUDA, "kstep_distributed_fill: %s: %s", what, cudaGetErrorString(err));
        return err == cudaSuccess;
    };
    if (st == PSIM_OK) st = tmp.alloc(&d_chunk, (size_t)std::max(m, 1));
    if (st == PSIM_OK) st = tmp.alloc(&d_rec, (size_t)std::max(m, 1));
    if (st == PSIM_OK) st = tmp.alloc(&d_dest, (size_t)std::max(m, 1));
    if (st == PSIM_OK) st = tmp.alloc(&d_ids, (size_t)std::max(m, 1));
    if (st == PSIM_OK) st = tmp.alloc(&d_counts, (size_t)3 * R);   // counts | offsets | cursors
    if (st == PSIM_OK) st = tmp.alloc(&d_bounds, (size_t)R + 1);
    if (st == PSIM_OK) {
        cuda_ok(cudaMemcpyAsync(d_chunk, parts + lo, sizeof(particle_t) * (size_t)m, cudaMemcpyHostToDevice, s), "upload");
        cuda_ok(cudaMemcpyAsync(d_bounds, bounds.data(), sizeof(int) * ((size_t)R + 1), cudaMemcpyHostToDevice, s), "bounds");
        cuda_ok(cudaMemsetAsync(d_counts, 0, sizeof(int) * 3 * (size_t)R, s), "memset");
        if (m > 0) kstep_route_count_kernel<<<(m + 255) / 256, 256, 0, s>>>(d_chunk, m, sim->bincnt, e->ts, R, d_bounds, d_dest, d_counts);
        ++sim->launches;
        cuda_ok(cudaMemcpyAsync(matrix.data() + (size_t)sim->rank * R, d_counts, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, s), "counts");
        cuda_ok(cudaStreamSynchronize(s), "sync");
    }
    lap("alloc + upload + route");
    // everybody learns the whole count matrix: matrix[src * R + dst]
    if (st == PSIM_OK) st = comm_allreduce_ints(sim, matrix.data(), R * R, s);
    lap("allreduce counts");
    long long n_recv = 0;
    if (st == PSIM_OK) {
        for (int r = 0; r < R; ++r) offs[(size_t)r + 1] = offs[(size_t)r] + matrix[(size_t)sim->rank * R + r];
        for (int src = 0; src < R; ++src) n_recv += matrix[(size_t)src * R + sim->rank];
        cuda_ok(cudaMemcpyAsync(d_counts + R, offs.data(), sizeof(int) * (size_t)R, cudaMemcpyHostToDevice, s), "offsets");
        if (m > 0) kstep_route_pack_kernel<<<(m + 255) / 256, 256, 0, s>>>(d_chunk, m, lo, d_dest, d_counts + R, d_counts + 2 * R, d_rec, d_ids);
        ++sim->launches;
        cuda_ok(cudaGetLastError(), "route");
    }
    if (st == PSIM_OK) st = tmp.alloc(&d_recv, (size_t)std::max<long long>(n_recv, 1));
    if (st == PSIM_OK) st = tmp.alloc(&d_recv_ids, (size_t)std::max<long long>(n_recv, 1));
    lap("pack + alloc recv");
    if (st == PSIM_OK) st = comm_alltoall_records(sim, d_rec, d_ids, offs.data(), d_recv, d_recv_ids, matrix.data(), sizeof(double4), s);
    lap("all-to-all");
    if (st == PSIM_OK && n_recv > 0) {
        kstep_fill_records_kernel<<<(unsigned)((n_recv + 255) / 256), 256, 0, s>>>(d_recv, d_recv_ids, (int)n_recv, sim->bincnt, e->ts, e->cap, e->ntx,
                                                                                 e->tr_begin, e->tr_end, e->tr_begin - 1, e->pos[0], e->vel[0],
                                                                                 e->sid[0], e->tcount, sim->d_err);
        ++sim->launches;
        cuda_ok(cudaGetLastError(), "fill");
    }
    if (st == PSIM_OK) cuda_ok(cudaStreamSynchronize(s), "sync");
    lap("fill stripes");
    tmp.release();
    lap("free staging");
    PSIM_TRY(st);
    bool unsuitable = false;
    st = kstep_finish_fill(sim, &unsuitable);
    lap("check + partition");
    return st;
}

// issue one launch (all owned rows) of nsub steps; slabs split it into a boundary and an interior launch
static int kstep_issue(psim_sim* sim, KstepEngine* e, int nsub, bool store) {
    cudaStream_t s = sim->stream;
    const int seq = e->seq++;
    if (e->log.size() >= (1u << 16)) e->log.erase(e->log.begin(), e->log.begin() + (1u << 15));   // (a caller that never synchronises)
    e->log.push_back({seq, nsub, e->parity, store, sim->steps_done});
    if (sim->nranks == 1) {
        PSIM_TRY(klaunch_rows(sim, e, e->parity, nsub, store, seq, 1, e->lrows, 1, false, s));
    } else if (sim->p2p) {
        // Peer-memory flavour (same choreography as the tiled engine, once per LAUNCH instead of once per step): the first and
        // last owned tile rows go first on the high-priority stream and store their facing bands straight into the neighbours'
        // ghost rows; a flag then tells the neighbours that my boundary rows of this launch are done.  My boundary rows may
        // start once both neighbours have flagged the previous launch.  Boundary and interior launches of one batch touch
        // disjoint tiles and run concurrently; each depends on BOTH launches of the previous batch.
        if (!e->ghost_fresh) {
            PSIM_TRY(kstep_exchange(sim, e->parity, s));
            e->ghost_fresh = true;
        }
        cudaStream_t sb = sim->comm_stream;
        const unsigned k = sim->p2p_steps & 1u, kp = k ^ 1u;
        if (sim->p2p_steps == 0) {
            PSIM_CUDA(cudaEventRecord(sim->ev_i[kp], s));
            PSIM_CUDA(cudaEventRecord(sim->ev_b[kp], s));
        }
        PSIM_CUDA(cudaStreamWaitEvent(sb, sim->ev_i[kp], 0));
        PSIM_TRY(comm_p2p_wait(sim, sb));
        if (e->lrows <= 2) PSIM_TRY(klaunch_rows(sim, e, e->parity, nsub, store, seq, 1, e->lrows, 1, true, sb));
        else PSIM_TRY(klaunch_rows(sim, e, e->parity, nsub, store, seq, 1, 2, e->lrows - 1, true, sb));
        PSIM_CUDA(cudaEventRecord(sim->ev_b[k], sb));
        PSIM_TRY(comm_p2p_signal(sim, sb));
        PSIM_CUDA(cudaStreamWaitEvent(s, sim->ev_b[kp], 0));
        if (e->lrows > 2) PSIM_TRY(klaunch_rows(sim, e, e->parity, nsub, store, seq, 2, e->lrows - 2, 1, false, s));
        PSIM_CUDA(cudaEventRecord(sim->ev_i[k], s));
        PSIM_CUDA(cudaStreamWaitEvent(s, sim->ev_b[k], 0));   // the handle's stream always covers the whole batch
    } else {
        // NCCL flavour: boundary rows first, their rows travel on the exchange stream while the interior rows are computed
        if (!e->ghost_fresh) {
            PSIM_TRY(kstep_exchange(sim, e->parity, s));
            e->ghost_fresh = true;
        }
        if (e->lrows <= 2) PSIM_TRY(klaunch_rows(sim, e, e->parity, nsub, store, seq, 1, e->lrows, 1, false, s));
        else PSIM_TRY(klaunch_rows(sim, e, e->parity, nsub, store, seq, 1, 2, e->lrows - 1, false, s));
        PSIM_CUDA(cudaEventRecord(sim->ev_boundary, s));
        PSIM_CUDA(cudaStreamWaitEvent(sim->comm_stream, sim->ev_boundary, 0));
        PSIM_TRY(kstep_exchange(sim, e->parity ^ 1, sim->comm_stream));
        PSIM_CUDA(cudaEventRecord(sim->ev_exchanged, sim->comm_stream));
        if (e->lrows > 2) PSIM_TRY(klaunch_rows(sim, e, e->parity, nsub, store, seq, 2, e->lrows - 2, 1, false, s));
        PSIM_CUDA(cudaStreamWaitEvent(s, sim->ev_exchanged, 0));
    }
    e->parity ^= 1;
    e->acc_valid = store;
    sim->steps_done += nsub;
    return PSIM_OK;
}

int kstep_step(psim_sim* sim, int nsteps, int flags) {
    KstepEngine* e = sim->kstep;
    if (sim->nranks > 1 && !sim->comm) return fail(PSIM_ERR_STATE, "slab %d/%d is not connected: call psim_comm_connect first", sim->rank, sim->nranks);
    const int kmax = (flags & PSIM_STEP_ACCEL_ALL) ? 1 : e->ksteps;
    int done = 0;
    while (done < nsteps) {
        const int k = std::min(kmax, nsteps - done);
        const bool store = (flags & PSIM_STEP_ACCEL_ALL) || (!(flags & PSIM_STEP_ACCEL_NONE) && done + k == nsteps);
        PSIM_TRY(kstep_issue(sim, e, k, store));
        done += k;
    }
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

// Called with the stream synchronised and the error words in sim->h_err.  No failed launch: forget the log.  A launch that
// exceeded the speed bound: rewind to its (untouched) input and replay it one step per launch -- the halo argument then
// allows a displacement of H - 1 cells per step -- followed by the launches that were skipped.  Returns PSIM_OK if the
// caller should synchronise and check again (`*replayed` set), or PSIM_ERR_CAPACITY; in that case *pending_steps >= 0 says
// that the state was rewound to the failed launch's input and how many enqueued steps are still owed.
int kstep_after_sync(psim_sim* sim, bool* replayed, int* pending_steps, bool* pending_store) {
    KstepEngine* e = sim->kstep;
    *replayed = false;
    *pending_steps = -1;
    *pending_store = false;
    const int failed = sim->h_err[8];
    if (failed == 0) {
        e->log.clear();
        return PSIM_OK;
    }
    const int flags = sim->h_err[0];
    const int seq = failed - 1;
    size_t at = 0;
    while (at < e->log.size() && e->log[at].seq != seq) ++at;
    const bool known = at < e->log.size();
    if ((flags & ~kErrSpeedBound) || !known || sim->nranks > 1 || e->log[at].nsub <= 1) {
        // capacity overflows, slabs (the neighbours have moved on) and a speed beyond even the one-step bound are not recoverable
        if (known) {   // the input of the failed launch is still intact: rewind to it; the caller may finish the batch on another engine
            e->parity = e->log[at].parity_in;
            sim->steps_done = e->log[at].steps_before;
            e->acc_valid = false;
            *pending_steps = 0;
            for (size_t r = at; r < e->log.size(); ++r) {
                *pending_steps += e->log[r].nsub;
                *pending_store = e->log[r].store;
            }
        }
        e->log.clear();
        return PSIM_ERR_CAPACITY;
    }
    std::vector<KLaunchRec> redo(e->log.begin() + at, e->log.end());
    e->log.clear();
    e->parity = redo[0].parity_in;
    sim->steps_done = redo[0].steps_before;
    PSIM_CUDA(cudaMemsetAsync(sim->d_err, 0, 9 * sizeof(int), sim->stream));
    ++e->recoveries;
    for (size_t r = 0; r < redo.size(); ++r) {
        if (r == 0) {
            for (int k = 0; k < redo[0].nsub; ++k) PSIM_TRY(kstep_issue(sim, e, 1, redo[0].store && k == redo[0].nsub - 1));
        } else {
            PSIM_TRY(kstep_issue(sim, e, redo[r].nsub, redo[r].store));
        }
    }
    PSIM_CUDA(cudaGetLastError());
    *replayed = true;
    return PSIM_OK;
}

int kstep_view(psim_sim* sim, SoAView* out) {
    KstepEngine* e = sim->kstep;
    cudaStream_t s = sim->stream;
    // scratch sized by what this slab can own (every owned stripe full), not by the global particle count
    const int capacity = (int)std::min<long long>((long long)sim->n_total, (long long)e->lrows * e->ntx * e->cap);
    if (!e->g_cursor || e->g_capacity < capacity) {
        e->gmem.release();
        const size_t c = (size_t)capacity + 2;
        PSIM_TRY(e->gmem.reserve(c * (6 * sizeof(double) + sizeof(int)) + 8 * 256 + 256));
        PSIM_TRY(e->gmem.alloc(&e->g.x, c));
        PSIM_TRY(e->gmem.alloc(&e->g.y, c));
        PSIM_TRY(e->gmem.alloc(&e->g.vx, c));
        PSIM_TRY(e->gmem.alloc(&e->g.vy, c));
        PSIM_TRY(e->gmem.alloc(&e->g.ax, c));
        PSIM_TRY(e->gmem.alloc(&e->g.ay, c));
        PSIM_TRY(e->gmem.alloc(&e->g.id, c));
        PSIM_TRY(e->gmem.alloc(&e->g_cursor, 1));
        e->g_capacity = capacity;
    }
    PSIM_CUDA(cudaMemsetAsync(e->g_cursor, 0, sizeof(int), s));
    const int p = e->parity;
    if (e->lrows * e->ntx > 0)
        kstep_gather_kernel<<<e->lrows * e->ntx, 128, 0, s>>>(e->pos[p], e->vel[p], e->sid[p], e->hdr[p], e->acc, e->cap, e->ntx,
                                                              e->acc_valid, e->g.x, e->g.y, e->g.vx, e->g.vy, e->g.ax, e->g.ay,
                                                              e->g.id, e->g_cursor);
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
int kstep_writeback(psim_sim* sim, particle_t* d_out, double2* d_xy) {
    KstepEngine* e = sim->kstep;
    const int p = e->parity;
    if (e->lrows * e->ntx > 0)
        kstep_writeback_kernel<<<e->lrows * e->ntx, 128, 0, sim->stream>>>(e->pos[p], e->vel[p], e->sid[p], e->hdr[p], e->acc, e->cap,
                                                                           e->ntx, e->acc_valid, d_out, d_xy);
    ++sim->launches;
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

// positions of the particles in the ghost rows' facing bands (the neighbours' boundary bands): neighbours for statistics
__global__ void __launch_bounds__(128) kstep_ghost_gather_kernel(const double2* __restrict__ pos, const int* __restrict__ hdr, int cap, int ntx,
                                                                 int row_lo, int row_hi, bool have_lo, bool have_hi, double* __restrict__ gx,
                                                                 double* __restrict__ gy, int capacity, int* __restrict__ cursor) {
    __shared__ int s_base;
    const bool hi = blockIdx.x >= (unsigned)ntx;
    if (hi ? !have_hi : !have_lo) return;
    const int lt = (hi ? row_hi : row_lo) * ntx + (int)(blockIdx.x % ntx);
    // lower ghost row = the lower neighbour's LAST row: its bottom band BR B BL (classes 4..6); upper ghost row: TL T TR (0..2)
    const int* h = hdr + (size_t)lt * kHdrInts;
    const int s0 = min(h[hi ? 0 : 4], cap), s1 = min(h[hi ? 3 : 7], cap), n = max(s1 - s0, 0);
    if (threadIdx.x == 0) s_base = n ? atomicAdd(cursor, n) : 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (s_base + i >= capacity) break;
        const double2 p = pos[(size_t)lt * cap + s0 + i];
        gx[s_base + i] = p.x;
        gy[s_base + i] = p.y;
    }
}

int kstep_ghost_capacity(psim_sim* sim) { return 2 * sim->kstep->ntx * sim->kstep->cap; }

int kstep_ghost_positions(psim_sim* sim, double* gx, double* gy, int capacity, int* count) {
    KstepEngine* e = sim->kstep;
    cudaStream_t s = sim->stream;
    *count = 0;
    if (sim->nranks == 1) return PSIM_OK;
    if (!e->ghost_fresh) {
        PSIM_TRY(kstep_exchange(sim, e->parity, s));
        e->ghost_fresh = true;
    }
    if (sim->p2p) PSIM_TRY(comm_p2p_wait(sim, s));   // the neighbours' stores into my ghost rows are complete
    int* d_cursor = nullptr;
    PSIM_CUDA(cudaMalloc(&d_cursor, sizeof(int)));
    cudaError_t err = cudaMemsetAsync(d_cursor, 0, sizeof(int), s);
    const int p = e->parity;
    if (err == cudaSuccess) {
        kstep_ghost_gather_kernel<<<2 * e->ntx, 128, 0, s>>>(e->pos[p], e->hdr[p], e->cap, e->ntx, 0, e->lrows + 1, sim->rank > 0,
                                                             sim->rank < sim->nranks - 1, gx, gy, capacity, d_cursor);
        ++sim->launches;
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpyAsync(count, d_cursor, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    cudaFree(d_cursor);
    if (err != cudaSuccess) return fail(PSIM_ERR_CUDA, "kstep_ghost_positions: %s", cudaGetErrorString(err));
    *count = std::min(*count, capacity);
    return PSIM_OK;
}

void kstep_destroy(psim_sim* sim) {
    KstepEngine* e = sim->kstep;
    if (!e) return;
    e->gmem.release();
    e->mem.release();
    delete e;
    sim->kstep = nullptr;
}

long long kstep_bytes(psim_sim* sim) { return sim->kstep ? (long long)(sim->kstep->mem.bytes + sim->kstep->gmem.bytes) : 0; }

void kstep_info(psim_sim* sim, psim_info_t* out) {
    KstepEngine* e = sim->kstep;
    out->tile_cells = e->ts;
    out->tiles_per_side = e->ntx;
    out->tile_capacity = e->cap;
    out->halo_cells = e->h;
    out->steps_per_launch = e->ksteps;
    out->region_capacity = e->nmax;
    out->recoveries = e->recoveries;
}

// accessors for psim_comm.cpp: the buffers shared with the neighbour slabs and the four byte ranges that make up one tile row
void kstep_shared_buffers(psim_sim* sim, char** parity0, char** parity1, size_t* bytes, int* ntx) {
    KstepEngine* e = sim->kstep;
    *parity0 = e->buf[0];
    *parity1 = e->buf[1];
    *bytes = e->buf_bytes;
    *ntx = e->ntx;
}

void kstep_row_ranges(psim_sim* sim, int parity, int lrow, char* ptr[4], size_t bytes[4]) {
    KstepEngine* e = sim->kstep;
    const size_t t0 = (size_t)lrow * e->ntx, s0 = t0 * e->cap;
    ptr[0] = reinterpret_cast<char*>(e->hdr[parity] + t0 * kHdrInts);
    bytes[0] = (size_t)e->ntx * kHdrInts * sizeof(int);
    ptr[1] = reinterpret_cast<char*>(e->pos[parity] + s0);
    bytes[1] = (size_t)e->ntx * e->cap * sizeof(double2);
    ptr[2] = reinterpret_cast<char*>(e->vel[parity] + s0);
    bytes[2] = bytes[1];
    ptr[3] = reinterpret_cast<char*>(e->sid[parity] + s0);
    bytes[3] = (size_t)e->ntx * e->cap * sizeof(int);
}

int kstep_owned_rows(psim_sim* sim) { return sim->kstep->lrows; }

}  // namespace psim
