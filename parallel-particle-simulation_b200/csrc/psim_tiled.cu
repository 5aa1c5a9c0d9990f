// psim_tiled.cu -- the "tiled" engine: persistent tile-resident particles, one fused kernel per step.
//
// Layout in HBM.  The box is cut into square tiles of TS x TS cutoff cells.  Every tile owns a stripe
// of CAP particle slots in five structure-of-arrays streams (x, y, vx, vy, id; plus ax, ay written
// only on steps whose accelerations are kept).  A tile's live particles occupy slots [0, count) and a
// particle keeps its slot for as long as it stays in the tile.  Steady-state HBM traffic is therefore
// read 32 B + write 32 B per particle-step plus a few percent of boundary lists: there is no global
// histogram, scan or scatter in the time loop.
//
// Per step one persistent kernel (tile_step_kernel); a CTA walks tiles blockIdx.x, +gridDim.x, ...
// with a two-stage shared-memory pipeline fed by TMA bulk copies (cp.async.bulk + mbarrier):
//   while tile k is computed, the bulk copies of tile k+1 (its 5 stripes, the 8 neighbour halo lists
//   and the 9 surrounding outboxes) are in flight and the list counts of tile k+2 are being fetched,
//   so no warp ever waits on a dependent chain of global loads.
// Compute on a staged tile:
//   A. ingest: particles that entered the tile last step (outbox records of the 3x3 tiles) are
//      appended; the one-cell apron is assembled from the neighbours' edge / corner halo lists (and
//      from outbox records that sit in the apron)
//   B. bin own + apron particles into a (TS+2)^2 table IN SHARED MEMORY: per cutoff cell a
//      population word and up to four particle indices (one atomicAdd + one 16-bit store per particle;
//      reference part3/gpu.cu:92-112 does this in global memory with 16 slots per cell)
//   C. force: every own particle reads the 9 words of its 3x3 neighbourhood (reference
//      part1/serial.cpp:102-117, 19-36), canonical summation order; cells with more than four
//      particles switch the tile to an exact all-pairs sweep
//   D. move + reflect (reference part1/serial.cpp:46-61) in registers
//   E. re-tile: stayers are written back to their own slot, leavers go to the tile's outbox and the
//      holes they leave are filled from the tail; particles in the tile's boundary cells are
//      appended to the edge / corner halo lists the neighbours read next step.
// Halo lists and outboxes ("exports") are double buffered by step parity: a step reads parity p and
// writes parity p^1, so no CTA ever reads what another CTA of the same launch writes.
//
// Slabs (precedent reference part2/mpi.cpp:258-270,296-365): a rank owns a
// contiguous range of tile rows plus one ghost tile row on each side that holds only exports.  The
// exports of one tile row are one contiguous byte range, so the halo exchange AND the particle
// migration between GPUs are a single send/receive of the first / last owned row per neighbour.
#include <algorithm>
#include <cstring>

#include "psim_force.cuh"
#include "psim_internal.h"
#include "psim_tiled.h"

namespace psim {

// ------------------------------------------------------------------------------------------
// compile-time tile configurations
// ------------------------------------------------------------------------------------------
// CAP slots per tile (mean population is 0.2*TS^2), HE / HC entries per edge / corner halo list,
// CO outbox records per tile of which the first CS are staged in shared memory by the pipeline (a
// tile that receives more from one neighbour reads the rest straight from global memory: this only
// happens in bursts, e.g. a column of the initial lattice that sits exactly on a tile boundary),
// THREADS per CTA (each thread owns PER = CAP/THREADS slots), CTAS resident per SM.
template <int TS> struct TileCfg;
// THREADS are the CONSUMER threads; every CTA has one more warp, the producer, that only feeds the pipeline.
// THREADS is a little above the MEAN population, so nearly every lane has a particle in the first pass;
// the second pass (PER = 2) only runs for the slots past THREADS.
template <> struct TileCfg<16> { static constexpr int CAP = 128,  HE = 24, HC = 8, CO = 64, CS = 4,  THREADS = 64,  CTAS = 8; };
template <> struct TileCfg<32> { static constexpr int CAP = 352,  HE = 32, HC = 8, CO = 64, CS = 4,  THREADS = 224, CTAS = 3; };
template <> struct TileCfg<64> { static constexpr int CAP = 1152, HE = 64, HC = 8, CO = 64, CS = 8,  THREADS = 832, CTAS = 1; };

struct __align__(16) OutRec {  // one migrating particle, 64 bytes
    double x, y, vx, vy, ax, ay;
    int id;
    int pad[3];
};
static_assert(sizeof(OutRec) == 64, "OutRec must be 64 bytes");

template <int TS> struct TileDims {
    using C = TileCfg<TS>;
    static constexpr int W = TS + 2, NC = W * W;
    static constexpr int HL = 4 * C::HE + 4 * C::HC;      // halo entries a tile exports / stages
    static constexpr int MAXH = HL + 32;                  // apron capacity (halo lists + apron outbox records)
    static constexpr int PTOT = C::CAP + MAXH;
    static constexpr int PER = (C::CAP + C::THREADS - 1) / C::THREADS;
    static constexpr int LMAX = C::CO;                    // leaver list capacity
    static_assert(C::CAP % 4 == 0 && PTOT % 2 == 0 && C::THREADS % 32 == 0, "alignment");
};

// one pipeline stage in shared memory (TMA destinations first, 16-byte aligned)
template <int TS> struct __align__(16) Stage {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    double x[D::PTOT], y[D::PTOT];   // own [0, n) then apron [CAP, CAP + n_apron)
    double vx[C::CAP], vy[C::CAP];
    int id[C::CAP];
    double2 halo[D::HL];             // raw halo lists of the 8 neighbours
    OutRec obox[9][C::CS];           // first CS records of the outboxes of the 3x3 tiles
};

template <int TS> struct __align__(16) TileSmem {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    Stage<TS> st[2];
    unsigned long long cell[D::NC];           // per cell: population | idx0 | idx1 | idx2 (16 bits each; atomics only)
    unsigned long long full[2];               // mbarriers: stage filled by TMA (producer -> consumers)
    unsigned long long empty[2];              // mbarriers: stage released (consumers -> producer)
    unsigned short pcell[D::PTOT];
    int cnts[2][20];                          // list counts of the staged tiles: [0] own, [1..8] halo, [9..17] outbox
    int leave[D::LMAX];                       // slots of the particles that leave this step
    int hole_dst[D::LMAX];                    // destination slot of the tail stayers that fill holes
    int hole[D::LMAX];
    int n_own, n_apron, n_leave, flags, overflow;
    int hw_leave, hw_halo, hw_tile, hw_apron;   // high-water marks over the tiles this CTA processed
    int hout[8];
};

__host__ __device__ inline int halo_offset(int list, int HE, int HC) { return list < 4 ? list * HE : 4 * HE + (list - 4) * HC; }
__host__ __device__ inline int halo_cap(int list, int HE, int HC) { return list < 4 ? HE : HC; }

struct TileParams {
    double *sx, *sy, *svx, *svy, *sax, *say;
    int* sid;
    int* tcount;
    const char* exp_in;
    char* exp_out;
    ExportLayout L;
    int ntx, nty;    // tiles per side (global)
    int tr_base;     // global tile row of local row 0
    int lrow0;       // first local tile row of this launch
    int ntiles;      // tiles of this launch (rows * ntx)
    int bincnt;
    double size;
    int* err;
};

__device__ __forceinline__ const char* row_ptr(const char* base, const ExportLayout& L, int lrow) {
    return base + (size_t)lrow * L.row_bytes;
}
__device__ __forceinline__ char* row_ptr(char* base, const ExportLayout& L, int lrow) {
    return base + (size_t)lrow * L.row_bytes;
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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the consumer threads only (id 1; id 0 is __syncthreads)
template <int N>
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tile index of this launch -> local row / column
struct TileCoord {
    int lr, tc, tr, lt;
};
__device__ __forceinline__ TileCoord tile_coord(const TileParams& P, int t) {
    TileCoord c;
    c.lr = P.lrow0 + t / P.ntx;
    c.tc = t % P.ntx;
    c.tr = P.tr_base + c.lr;
    c.lt = c.lr * P.ntx + c.tc;
    return c;
}

// list count j of tile t: j = 0 own population, 1..8 apron sources, 9..17 outboxes of the 3x3 tiles
template <int TS>
__device__ __forceinline__ int load_count(const TileParams& P, int t, int j) {
    using C = TileCfg<TS>;
    const TileCoord c = tile_coord(P, t);
    if (j == 0) return min(P.tcount[c.lt], C::CAP);
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
    return min(ec[list], list == 8 ? C::CO : halo_cap(list, C::HE, C::HC));
}

// Issue the bulk copies of tile t into `st` (one warp; lane j owns copy j).
//   lanes 0-3: x y vx vy stripes, lane 4: id stripe, lanes 5-12: apron sources, lanes 13-21: outboxes
template <int TS>
__device__ __forceinline__ void issue_tile_loads(const TileParams& P, int t, Stage<TS>& st, const int* cnt,
                                                 unsigned long long* bar, int lane) {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    const TileCoord c = tile_coord(P, t);
    const size_t gbase = (size_t)c.lt * C::CAP;
    const void* src = nullptr;
    void* dst = nullptr;
    unsigned bytes = 0;
    if (lane < 5) {
        const int n = cnt[0];
        if (lane < 4) {
            bytes = (unsigned)((n + 1) >> 1) * 16u;
            src = (lane == 0 ? P.sx : lane == 1 ? P.sy : lane == 2 ? P.svx : P.svy) + gbase;
            dst = lane == 0 ? st.x : lane == 1 ? st.y : lane == 2 ? st.vx : st.vy;
        } else {
            bytes = (unsigned)((n + 3) >> 2) * 16u;
            src = P.sid + gbase;
            dst = st.id;
        }
    } else if (lane < 13) {
        const int k = lane - 5;
        int dr, dc, list;
        halo_source(k, dr, dc, list);
        bytes = (unsigned)cnt[1 + k] * 16u;
        if (bytes) {
            src = reinterpret_cast<const double2*>(row_ptr(P.exp_in, P.L, c.lr + dr) + P.L.off_hxy) +
                  (size_t)(c.tc + dc) * D::HL + halo_offset(list, C::HE, C::HC);
            dst = st.halo + halo_offset(k, C::HE, C::HC);
        }
    } else if (lane < 22) {
        const int nb = lane - 13;
        const int dr = nb / 3 - 1, dc = nb % 3 - 1;
        bytes = (unsigned)min(cnt[9 + nb], C::CS) * (unsigned)sizeof(OutRec);
        if (bytes) {
            src = reinterpret_cast<const OutRec*>(row_ptr(P.exp_in, P.L, c.lr + dr) + P.L.off_obox) + (size_t)(c.tc + dc) * C::CO;
            dst = st.obox[nb];
        }
    }
    unsigned total = bytes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) mbar_arrive_expect_tx(bar, total);
    __syncwarp();
    if (bytes) tma_load_1d(dst, src, bytes, bar);
}

// ------------------------------------------------------------------------------------------
// the per-step kernel
// ------------------------------------------------------------------------------------------
template <int TS, bool kStoreAcc>
__global__ void __launch_bounds__(TileCfg<TS>::THREADS + 32, TileCfg<TS>::CTAS) tile_step_kernel(const TileParams P) {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    constexpr int T = C::THREADS, CAP = C::CAP, W = D::W, NC = D::NC, PER = D::PER;
    constexpr int HE = C::HE, HC = C::HC, CO = C::CO, LMAX = D::LMAX;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem<TS>& S = *reinterpret_cast<TileSmem<TS>*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int first = blockIdx.x;
    if (first >= P.ntiles) return;

    // ---- prologue ---------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        mbar_init(&S.empty[0], 1);
        mbar_init(&S.empty[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.n_leave = 0;
        S.flags = 0;
        S.overflow = 0;
        S.hw_leave = S.hw_halo = S.hw_tile = S.hw_apron = 0;
    }
    if (tid < 8) S.hout[tid] = 0;
    for (int c = tid; c < NC; c += T + 32) S.cell[c] = 0ull;
    __syncthreads();

    // ---- producer warp: runs up to two tiles ahead of the consumers ----------------------------------
    if (warp == T / 32) {
        for (int j = 0;; ++j) {
            const int t = first + j * G;
            if (t >= P.ntiles) break;
            const int c = lane < 18 ? load_count<TS>(P, t, lane) : 0;
            if (j >= 2) mbar_wait(&S.empty[j & 1], (unsigned)(((j >> 1) - 1) & 1));  // tile j-2 has left the stage
            if (lane < 18) S.cnts[j & 1][lane] = c;
            __syncwarp();
            issue_tile_loads<TS>(P, t, S.st[j & 1], S.cnts[j & 1], &S.full[j & 1], lane);
        }
        return;
    }

    // ---- consumers ----------------------------------------------------------------------------------
    for (int it = 0;; ++it) {
        const int t = first + it * G;
        if (t >= P.ntiles) break;
        Stage<TS>& st = S.st[it & 1];
        const int* cnt = S.cnts[it & 1];
        const TileCoord tc_ = tile_coord(P, t);
        const int tr = tc_.tr, tc = tc_.tc, lt = tc_.lt, lr = tc_.lr;
        const int r0 = tr * TS, c0 = tc * TS;
        const size_t gbase = (size_t)lt * CAP;

        mbar_wait(&S.full[it & 1], (unsigned)((it >> 1) & 1));  // this tile's bytes (and counts) have landed

        const int n_own0 = cnt[0];

        // ---- A: ingest newcomers and assemble the apron ---------------------------------------------
        int hoff[9];
        hoff[0] = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) hoff[k + 1] = hoff[k] + cnt[1 + k];
        if (warp == 0) {
            // outbox records of the 3x3 tiles: records now in my tile are appended to my stripe, records in my
            // apron ring become apron particles (after the halo-list entries).  One lane per staged record.
            const unsigned lt_mask = (1u << lane) - 1u;
            int n_own = n_own0, n_ap = hoff[8], flags = 0;
            auto take = [&](bool valid, const OutRec* rec) {
                bool mine = false, apron = false;
                double x = 0, y = 0;
                if (valid) {
                    x = rec->x;
                    y = rec->y;
                    const int row_c = axis_cell(x, P.bincnt), col_c = axis_cell(y, P.bincnt);
                    mine = row_c / TS == tr && col_c / TS == tc;
                    apron = !mine && row_c >= r0 - 1 && row_c <= r0 + TS && col_c >= c0 - 1 && col_c <= c0 + TS;
                }
                const unsigned mm = __ballot_sync(0xffffffffu, mine), am = __ballot_sync(0xffffffffu, apron);
                if (mine) {
                    const int d = n_own + __popc(mm & lt_mask);
                    if (d < CAP) {
                        st.x[d] = x;
                        st.y[d] = y;
                        st.vx[d] = rec->vx;
                        st.vy[d] = rec->vy;
                        st.id[d] = rec->id;
                    }
                }
                if (apron) {
                    const int hh = n_ap + __popc(am & lt_mask);
                    if (hh < D::MAXH) {
                        st.x[CAP + hh] = x;
                        st.y[CAP + hh] = y;
                    }
                }
                n_own += __popc(mm);
                n_ap += __popc(am);
            };
            int total_out = 0, beyond = 0;
#pragma unroll
            for (int nb = 0; nb < 9; ++nb) {
                total_out += cnt[9 + nb];
                beyond |= cnt[9 + nb] > C::CS;
            }
            if (total_out > 0) {
#pragma unroll 1
                for (int q0 = 0; q0 < 9 * C::CS; q0 += 32) {
                    const int q = q0 + lane, nb = q / C::CS, e = q % C::CS;
                    const bool valid = q < 9 * C::CS && e < cnt[9 + min(nb, 8)];
                    take(valid, &st.obox[min(nb, 8)][e]);
                }
                if (beyond) {
                    // bursts only: records past the staged CS are read from the neighbour's outbox in global memory
#pragma unroll 1
                    for (int nb = 0; nb < 9; ++nb) {
                        const int c = cnt[9 + nb];
                        if (c <= C::CS) continue;
                        const OutRec* grec = reinterpret_cast<const OutRec*>(row_ptr(P.exp_in, P.L, lr + nb / 3 - 1) + P.L.off_obox) +
                                             (size_t)(tc + nb % 3 - 1) * CO;
                        for (int e0 = C::CS; e0 < c; e0 += 32) take(e0 + lane < c, grec + e0 + lane);
                    }
                }
            }
            if (n_own > CAP) { flags |= kErrTileOverflow; n_own = CAP; }
            if (n_ap > D::MAXH) { flags |= kErrSmemOverflow; n_ap = D::MAXH; }
            if (lane == 0) {
                S.n_own = n_own;
                S.n_apron = n_ap;
                if (flags) atomicOr(&S.flags, flags);
            }
        } else {
            // halo lists -> contiguous apron positions
            for (int q = tid - 32; q < hoff[8]; q += T - 32) {
                int k = 0;
#pragma unroll
                for (int j = 1; j < 8; ++j) k += q >= hoff[j];
                const double2 v = st.halo[halo_offset(k, HE, HC) + (q - hoff[k])];
                st.x[CAP + q] = v.x;
                st.y[CAP + q] = v.y;
            }
        }
        consumer_sync<T>();
        const int n_own = S.n_own, n_apron = S.n_apron;

        // ---- B: bin own + apron particles into the cell table -----------------------------------------
        for (int q = tid; q < n_own + n_apron; q += T) {
            const int i = q < n_own ? q : CAP + (q - n_own);
            const int lrow = axis_cell(st.x[i], P.bincnt) - r0 + 1, lcol = axis_cell(st.y[i], P.bincnt) - c0 + 1;
            unsigned short cell = 0xFFFFu;
            if (lrow >= 0 && lrow < W && lcol >= 0 && lcol < W) {
                cell = (unsigned short)(lrow * W + lcol);
                // every modification of the word is atomic, so count and indices never tear
                unsigned* w32 = reinterpret_cast<unsigned*>(&S.cell[cell]);
                const unsigned slot = atomicAdd(w32, 1u) & 0xFFFFu;
                if (slot < 3u) atomicOr(w32 + ((slot + 1u) >> 1), (unsigned)i << (16u * ((slot + 1u) & 1u)));
                else S.overflow = 1;
            }
            S.pcell[i] = cell;
        }
        consumer_sync<T>();
        const bool overflow = S.overflow != 0;

        // ---- C + D: force over the 3x3 neighbourhood, move; new state stays in registers --------------
        // new positions stay in registers (the old ones are still being read by other threads); new velocities
        // go back to the stage in place (only the owner reads them); accelerations are kept only when stored
        double nx[PER], ny[PER], nax[kStoreAcc ? PER : 1], nay[kStoreAcc ? PER : 1];
        int nrc[PER];  // new cell row << 16 | new cell column
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int i = r * T + tid;
            nrc[r] = 0;
            nx[r] = ny[r] = 0.0;
            if (i < n_own) {
                const double xi = st.x[i], yi = st.y[i];
                const int cell = S.pcell[i];
                double ax, ay;
                int nbc;
                const int lrow = cell / W, lcol = cell - lrow * W;
                // pass 0: collect the indices of the (at most eight) other particles in the 3x3 cells into two
                // registers; the distance loop below then runs over real candidates only, not over cells
                unsigned long long cand_lo = 0ull, cand_hi = 0ull;
                int ncand = 0;
                auto push = [&](unsigned j) {
                    if (j == (unsigned)i) return;
                    if (ncand < 4) cand_lo |= (unsigned long long)j << (16 * ncand);
                    else if (ncand < 8) cand_hi |= (unsigned long long)j << (16 * (ncand - 4));
                    ++ncand;
                };
                if (!overflow) {
#pragma unroll
                    for (int dr = -1; dr <= 1; ++dr) {
#pragma unroll
                        for (int dc = -1; dc <= 1; ++dc) {
                            const unsigned long long w = S.cell[cell + dr * W + dc];
                            const unsigned c = (unsigned)w & 0xFFFFu;
                            if (c > 0u) {
                                push((unsigned)(w >> 16) & 0xFFFFu);
                                if (c > 1u) push((unsigned)(w >> 32) & 0xFFFFu);
                                if (c > 2u) push((unsigned)(w >> 48));
                            }
                        }
                    }
                }
                const bool sweep = overflow || ncand > 8;
                auto visit = [&](auto&& f) {
                    if (!sweep) {
#pragma unroll 1
                        for (int k = 0; k < ncand; ++k) {
                            const int j = (int)(((k < 4 ? cand_lo : cand_hi) >> (16 * (k & 3))) & 0xFFFFull);
                            f(st.x[j], st.y[j], j);
                        }
                    } else {
                        // a cell near this particle holds more than three particles: exact sweep over everything staged
#pragma unroll 1
                        for (int q = 0; q < n_own + n_apron; ++q) {
                            const int j = q < n_own ? q : CAP + (q - n_own);
                            const int cj = S.pcell[j];
                            if (cj == 0xFFFF) continue;
                            const int dr = cj / W - lrow, dc = cj % W - lcol;
                            if (dr < -1 || dr > 1 || dc < -1 || dc > 1) continue;
                            f(st.x[j], st.y[j], j);
                        }
                    }
                };
                auto rank_of = [&](double, double, int j) {
                    const int cj = S.pcell[j];
                    return visit_rank(cj / W - lrow, cj % W - lcol);
                };
                accumulate_force(xi, yi, visit, rank_of, ax, ay, nbc);
                double x = xi, y = yi, vx = st.vx[i], vy = st.vy[i];
                move_particle(x, y, vx, vy, ax, ay, P.size);
                nx[r] = x; ny[r] = y;
                st.vx[i] = vx;
                st.vy[i] = vy;
                if (kStoreAcc) { nax[r] = ax; nay[r] = ay; }
                const int nrow = axis_cell(x, P.bincnt), ncol = axis_cell(y, P.bincnt);
                nrc[r] = (nrow << 16) | ncol;
                const int er = nrow - r0, ec = ncol - c0;
                if (er < 0 || er >= TS || ec < 0 || ec >= TS) {
                    const int k = atomicAdd(&S.n_leave, 1);
                    if (k < LMAX) S.leave[k] = i;
                }
            }
        }
        consumer_sync<T>();

        // ---- E: re-tile ---------------------------------------------------------------------------------
        const int n_leave_raw = S.n_leave;
        const int n_leave = min(n_leave_raw, LMAX);
        const int new_count = n_own - n_leave_raw;
        if (n_leave > 0) {
            // holes (leaver slots below new_count) are filled by the stayers at or above new_count, both in
            // ascending slot order: deterministic regardless of the arrival order in S.leave
            if (warp == 0) {
                for (int e = lane; e < n_leave; e += 32) {
                    const int s = S.leave[e];
                    if (s < new_count) {
                        int rank = 0;
                        for (int f = 0; f < n_leave; ++f) rank += S.leave[f] < s;
                        S.hole[rank] = s;  // rank among ALL leavers below s == rank among holes (holes are the smallest leaver slots)
                    }
                }
                __syncwarp();
                for (int m = new_count + lane; m < n_own; m += 32) {
                    bool is_leaver = false;
                    int leavers_below = 0;
                    for (int f = 0; f < n_leave; ++f) {
                        const int s = S.leave[f];
                        is_leaver |= s == m;
                        leavers_below += (s >= new_count && s < m);
                    }
                    if (m - new_count < LMAX) S.hole_dst[m - new_count] = is_leaver ? -1 : S.hole[(m - new_count) - leavers_below];
                }
            }
            consumer_sync<T>();
        }

        char* orow = row_ptr(P.exp_out, P.L, lr);
        double2* ohxy = reinterpret_cast<double2*>(orow + P.L.off_hxy) + (size_t)tc * D::HL;
        OutRec* oobox = reinterpret_cast<OutRec*>(orow + P.L.off_obox) + (size_t)tc * CO;
        int tflags = 0;
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int i = r * T + tid;
            if (i < n_own) {
                const int nrow = nrc[r] >> 16, ncol = nrc[r] & 0xFFFF;
                const int er = nrow - r0, ec = ncol - c0;
                const bool stay = er >= 0 && er < TS && ec >= 0 && ec < TS;
                const double nvx = st.vx[i], nvy = st.vy[i];
                if (stay) {
                    int d = i;
                    if (i >= new_count) d = (i - new_count < LMAX) ? S.hole_dst[i - new_count] : -1;
                    if (d >= 0) {
                        P.sx[gbase + d] = nx[r];
                        P.sy[gbase + d] = ny[r];
                        P.svx[gbase + d] = nvx;
                        P.svy[gbase + d] = nvy;
                        if (kStoreAcc) {
                            P.sax[gbase + d] = nax[kStoreAcc ? r : 0];
                            P.say[gbase + d] = nay[kStoreAcc ? r : 0];
                        }
                        if (d != i || i >= n_own0) P.sid[gbase + d] = st.id[i];
                    }
                    const bool n_ = er == 0, s_ = er == TS - 1, w_ = ec == 0, e_ = ec == TS - 1;
                    if (n_ | s_ | w_ | e_) {
                        const double2 q = make_double2(nx[r], ny[r]);
                        auto put = [&](int list) {
                            const int idx = atomicAdd(&S.hout[list], 1);
                            if (idx < halo_cap(list, HE, HC)) ohxy[halo_offset(list, HE, HC) + idx] = q;
                        };
                        // (with TS >= 3 a cell is on at most one of N/S and one of W/E)
                        if (n_ | s_) put(n_ ? 0 : 1);
                        if (w_ | e_) put(w_ ? 2 : 3);
                        if ((n_ | s_) && (w_ | e_)) put(4 + (s_ ? 2 : 0) + (e_ ? 1 : 0));
                    }
                } else {
                    int rank = 0;
                    for (int f = 0; f < n_leave; ++f) rank += S.leave[f] < i;
                    if (rank < CO) {
                        OutRec rec;
                        rec.x = nx[r]; rec.y = ny[r]; rec.vx = nvx; rec.vy = nvy;
                        rec.ax = kStoreAcc ? nax[kStoreAcc ? r : 0] : 0.0;
                        rec.ay = kStoreAcc ? nay[kStoreAcc ? r : 0] : 0.0;
                        rec.id = st.id[i];
                        rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
                        oobox[rank] = rec;
                    }
                    const int dtr = nrow / TS - tr, dtc = ncol / TS - tc;
                    if (dtr < -1 || dtr > 1 || dtc < -1 || dtc > 1) tflags |= kErrLostParticle;
                }
            }
        }
        if (tflags) atomicOr(&S.flags, tflags);
        fence_proxy_async();  // order this iteration's generic accesses to the stage before the next bulk copies
        consumer_sync<T>();
        if (tid == 0) mbar_arrive(&S.empty[it & 1]);  // the producer may refill this stage

        // counts out, reset the per-tile scratch for the next iteration
        if (tid < 9) {
            int* ec = reinterpret_cast<int*>(orow + P.L.off_cnt) + (size_t)tc * 16;
            if (tid < 8) {
                const int c = S.hout[tid], cap = halo_cap(tid, HE, HC);
                if (c > cap) atomicOr(&S.flags, kErrHaloOverflow);
                ec[tid] = min(c, cap);
                S.hout[tid] = 0;
                if (tid < 4) atomicMax(&S.hw_halo, c);
            } else {
                if (n_leave_raw > CO) atomicOr(&S.flags, kErrOutboxOverflow);
                ec[8] = min(n_leave_raw, CO);
                P.tcount[lt] = max(new_count, 0);
                S.n_leave = 0;
                S.overflow = 0;
                S.hw_leave = max(S.hw_leave, n_leave_raw);
                S.hw_tile = max(S.hw_tile, n_own);
                S.hw_apron = max(S.hw_apron, n_apron);
            }
        }
        for (int c = tid; c < NC; c += T) S.cell[c] = 0ull;
        consumer_sync<T>();
    }
    if (tid == 0) {
        if (S.flags) atomicOr(P.err, S.flags);
        atomicMax(P.err + 1, S.hw_leave);
        atomicMax(P.err + 2, S.hw_halo);
        atomicMax(P.err + 3, S.hw_tile);
        atomicMax(P.err + 4, S.hw_apron);
    }
}

// ------------------------------------------------------------------------------------------
// initial tiling
// ------------------------------------------------------------------------------------------
// one thread per input record: owned rows only; arrival order inside a tile is arbitrary, which is
// harmless because forces are summed in a canonical order.
__global__ void __launch_bounds__(256) tile_fill_kernel(const particle_t* __restrict__ p, int n, int id0, int bincnt,
                                                        int ts, int cap, int ntx, int tr_begin, int tr_end,
                                                        int tr_base, double* __restrict__ sx, double* __restrict__ sy,
                                                        double* __restrict__ svx, double* __restrict__ svy,
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
    sx[d] = a.x;
    sy[d] = a.y;
    svx[d] = b.x;
    svy[d] = b.y;
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
        const double x = P.sx[gbase + i], y = P.sy[gbase + i];
        const int er = axis_cell(x, P.bincnt) - r0, ec = axis_cell(y, P.bincnt) - c0;
        const bool n_ = er == 0, s_ = er == TS - 1, w_ = ec == 0, e_ = ec == TS - 1;
        if (n_ | s_ | w_ | e_) {
            const double2 q = make_double2(x, y);
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
__global__ void __launch_bounds__(128) tile_gather_kernel(const TileParams P, int ts, int cap, int co, int tr_begin, int tr_end,
                                                          bool have_acc, double* __restrict__ gx, double* __restrict__ gy,
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
            const int dtr = axis_cell(ob[e].x, P.bincnt) / ts;
            if (dtr >= tr_begin && dtr < tr_end) s_take[k++] = e;
        }
        s_ntake = k;
        s_base = (n_tile + k) ? atomicAdd(cursor, n_tile + k) : 0;
    }
    __syncthreads();
    const int base = s_base;
    const size_t gbase = (size_t)lt * cap;
    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
        gx[base + i] = P.sx[gbase + i];
        gy[base + i] = P.sy[gbase + i];
        gvx[base + i] = P.svx[gbase + i];
        gvy[base + i] = P.svy[gbase + i];
        gax[base + i] = have_acc ? P.sax[gbase + i] : 0.0;
        gay[base + i] = have_acc ? P.say[gbase + i] : 0.0;
        gid[base + i] = P.sid[gbase + i];
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
    double *sx = nullptr, *sy = nullptr, *svx = nullptr, *svy = nullptr, *sax = nullptr, *say = nullptr;
    int* sid = nullptr;
    int* tcount = nullptr;
    char* exports[2] = {nullptr, nullptr};
    size_t export_bytes = 0;  // one parity
    int parity = 0;           // exports[parity] is what the next step reads
    bool acc_valid = false;
    bool ghost_fresh = false;  // ghost rows hold the neighbours' exports of the current parity
    // gather scratch
    DeviceArena gmem;
    SoAView g{};
    int* g_cursor = nullptr;
    int g_capacity = 0;
};

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
    P.sx = e->sx; P.sy = e->sy; P.svx = e->svx; P.svy = e->svy; P.sax = e->sax; P.say = e->say;
    P.sid = e->sid;
    P.tcount = e->tcount;
    P.exp_in = e->exports[parity_in];
    P.exp_out = e->exports[parity_in ^ 1];
    P.L = e->L;
    P.ntx = e->ntx;
    P.nty = e->ntx;
    P.tr_base = e->tr_begin - 1;
    P.lrow0 = 1;
    P.ntiles = e->lrows * e->ntx;
    P.bincnt = sim->bincnt;
    P.size = sim->size;
    P.err = sim->d_err;
    return P;
}

template <int TS>
static int launch_step(psim_sim* sim, TiledEngine* e, int parity_in, bool store_acc, int lrow0, int nrows, cudaStream_t s) {
    if (nrows <= 0) return PSIM_OK;
    TileParams P = make_params(sim, e, parity_in);
    P.lrow0 = lrow0;
    P.ntiles = nrows * e->ntx;
    const int grid = std::min(P.ntiles, e->sms * e->ctas_per_sm);
    if (store_acc)
        tile_step_kernel<TS, true><<<grid, TileCfg<TS>::THREADS + 32, sizeof(TileSmem<TS>), s>>>(P);
    else
        tile_step_kernel<TS, false><<<grid, TileCfg<TS>::THREADS + 32, sizeof(TileSmem<TS>), s>>>(P);
    ++sim->launches;
    return PSIM_OK;
}

static int launch_step_ts(psim_sim* sim, TiledEngine* e, int parity_in, bool store_acc, int lrow0, int nrows, cudaStream_t s) {
    switch (e->ts) {
        case 16: return launch_step<16>(sim, e, parity_in, store_acc, lrow0, nrows, s);
        case 32: return launch_step<32>(sim, e, parity_in, store_acc, lrow0, nrows, s);
        case 64: return launch_step<64>(sim, e, parity_in, store_acc, lrow0, nrows, s);
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
    PSIM_CUDA(cudaFuncSetAttribute(tile_step_kernel<TS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    PSIM_CUDA(cudaFuncSetAttribute(tile_step_kernel<TS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    PSIM_CUDA(cudaFuncSetAttribute(tile_step_kernel<TS, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PSIM_CUDA(cudaFuncSetAttribute(tile_step_kernel<TS, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    PSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_step_kernel<TS, false>, TileCfg<TS>::THREADS + 32, e->smem));
    e->ctas_per_sm = std::max(1, per_sm);
    return PSIM_OK;
}

int tiled_tile_rows(int bincnt, int ts) { return (bincnt + ts - 1) / ts; }

void tiled_slab_rows(int ntx, int rank, int nranks, int* begin, int* end) {
    // contiguous tile rows, remainder spread over the first ranks
    const int q = ntx / nranks, r = ntx % nranks;
    *begin = rank * q + std::min(rank, r);
    *end = *begin + q + (rank < r ? 1 : 0);
}

int tiled_create(psim_sim* sim, const psim_config* cfg, const particle_t* parts, int n, bool parts_on_device,
                 bool* unsuitable) {
    *unsuitable = false;
    cudaStream_t s = sim->stream;
    int ts = cfg->tile_cells;
    if (ts == 0) {
        const long long cells = (long long)sim->bincnt * sim->bincnt;
        ts = cells >= 32ll * 32 * 148 * 4 ? 32 : 16;
    }
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
    PSIM_TRY(e->mem.alloc(&e->sx, slots));
    PSIM_TRY(e->mem.alloc(&e->sy, slots));
    PSIM_TRY(e->mem.alloc(&e->svx, slots));
    PSIM_TRY(e->mem.alloc(&e->svy, slots));
    PSIM_TRY(e->mem.alloc(&e->sax, slots));
    PSIM_TRY(e->mem.alloc(&e->say, slots));
    PSIM_TRY(e->mem.alloc(&e->sid, slots));
    PSIM_TRY(e->mem.alloc(&e->tcount, (size_t)e->lrows_alloc * e->ntx));
    PSIM_TRY(e->mem.alloc(&e->exports[0], e->export_bytes));
    PSIM_TRY(e->mem.alloc(&e->exports[1], e->export_bytes));
    PSIM_CUDA(cudaMemsetAsync(e->tcount, 0, sizeof(int) * (size_t)e->lrows_alloc * e->ntx, s));
    PSIM_CUDA(cudaMemsetAsync(e->exports[0], 0, e->export_bytes, s));
    PSIM_CUDA(cudaMemsetAsync(e->exports[1], 0, e->export_bytes, s));
    PSIM_CUDA(cudaMemsetAsync(e->sax, 0, sizeof(double) * slots, s));
    PSIM_CUDA(cudaMemsetAsync(e->say, 0, sizeof(double) * slots, s));
    // fill: device input is read in place; host input is streamed through a bounded staging buffer
    // (a slab keeps only its own rows, so no rank ever holds the whole array on the device)
    {
        DeviceArena stage;
        particle_t* d_stage = nullptr;
        const int chunk = parts_on_device ? n : std::min(n, 8 << 20);
        if (!parts_on_device && n > 0) PSIM_TRY(stage.alloc(&d_stage, (size_t)chunk));
        for (int off = 0; off < n; off += chunk) {
            const int m = std::min(chunk, n - off);
            const particle_t* src = parts + off;
            if (!parts_on_device) {
                PSIM_CUDA(cudaMemcpyAsync(d_stage, parts + off, sizeof(particle_t) * (size_t)m, cudaMemcpyHostToDevice, s));
                src = d_stage;
            }
            tile_fill_kernel<<<(m + 255) / 256, 256, 0, s>>>(src, m, off, sim->bincnt, ts, e->cap, e->ntx, e->tr_begin,
                                                             e->tr_end, e->tr_begin - 1, e->sx, e->sy, e->svx, e->svy,
                                                             e->sid, e->tcount, sim->d_err);
            ++sim->launches;
            if (!parts_on_device) PSIM_CUDA(cudaStreamSynchronize(s));
        }
        PSIM_CUDA(cudaGetLastError());
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
    TileParams P = make_params(sim, e, 1);  // writes exports[0]
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
    for (int step = 0; step < nsteps; ++step) {
        const bool store = (flags & PSIM_STEP_ACCEL_ALL) || (!(flags & PSIM_STEP_ACCEL_NONE) && step == nsteps - 1);
        if (sim->nranks > 1 && !e->ghost_fresh) {
            PSIM_TRY(tiled_exchange(sim, e->parity, s));
            e->ghost_fresh = true;
        }
        PSIM_TRY(launch_step_ts(sim, e, e->parity, store, 1, e->lrows, s));
        e->parity ^= 1;
        if (sim->nranks > 1) PSIM_TRY(tiled_exchange(sim, e->parity, s));
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
    TileParams P = make_params(sim, e, e->parity);
    tile_gather_kernel<<<e->lrows_alloc * e->ntx, 128, 0, s>>>(P, e->ts, e->cap, e->co, e->tr_begin, e->tr_end, e->acc_valid,
                                                               e->g.x, e->g.y, e->g.vx, e->g.vy, e->g.ax, e->g.ay, e->g.id,
                                                               e->g_cursor);
    ++sim->launches;
    PSIM_CUDA(cudaGetLastError());
    int n = 0;
    PSIM_CUDA(cudaMemcpyAsync(&n, e->g_cursor, sizeof(int), cudaMemcpyDeviceToHost, s));
    PSIM_CUDA(cudaStreamSynchronize(s));
    *out = e->g;
    out->n = n;
    return PSIM_OK;
}

void tiled_destroy(psim_sim* sim) {
    TiledEngine* e = sim->tiled;
    if (!e) return;
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

// accessors for psim_comm.cpp
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
