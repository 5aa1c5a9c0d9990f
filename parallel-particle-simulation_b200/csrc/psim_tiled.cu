// psim_tiled.cu -- the "tiled" engine: persistent tile-resident particles, one fused kernel per step.
//
// Layout in HBM.  The box is cut into square tiles of TS x TS cutoff cells.  Every tile owns CAP
// particle slots in five structure-of-arrays streams (x, y, vx, vy, id; plus ax, ay written only on
// steps whose accelerations are kept).  A tile's live particles occupy slots [0, count) of its
// stripe, so one CTA streams its tile with fully coalesced 8-byte-per-lane loads and stores and the
// steady-state HBM traffic is read 32 B + write 32 B per particle-step plus a few percent of
// boundary lists -- there is no global histogram, scan or scatter in the time loop.
//
// Per step, CTA = tile (kernel tile_step_kernel):
//   A. load own particles; ingest particles that entered the tile during the previous step from the
//      9 surrounding "outboxes"; load the one-cell apron around the tile from the neighbours'
//      edge / corner "halo lists" (and from outbox entries that sit in the apron)
//   B. bin own + apron particles into (TS+2)^2 cutoff cells IN SHARED MEMORY (atomic count,
//      block scan, index scatter) -- reference part3/gpu.cu:92-112 does this in global memory
//   C. force: every own particle walks its 3x3 cells = three contiguous shared-memory index
//      ranges (reference part1/serial.cpp:102-117, 19-36), canonical summation order
//   D. move + reflect (reference part1/serial.cpp:46-61) in registers
//   E. re-tile: particles still in the tile are compacted back into the tile's stripe (in place);
//      leavers go to the tile's outbox; particles now in the tile's boundary cells are appended to
//      the edge / corner halo lists the neighbours will read next step.
// Halo lists and outboxes ("exports") are double buffered by step parity, so a step reads parity p
// and writes parity p^1 and no CTA ever reads data another CTA of the same launch writes.
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
template <int TS> struct TileCfg;
template <> struct TileCfg<16> { static constexpr int CAP = 128,  HE = 24, HC = 8, CO = 16, THREADS = 64;  };
template <> struct TileCfg<32> { static constexpr int CAP = 384,  HE = 40, HC = 8, CO = 32, THREADS = 256; };
template <> struct TileCfg<64> { static constexpr int CAP = 1152, HE = 64, HC = 8, CO = 48, THREADS = 512; };

template <int TS> struct TileDims {
    using C = TileCfg<TS>;
    static constexpr int W = TS + 2, NC = W * W;
    static constexpr int MAXH = 4 * C::HE + 4 * C::HC + 32;
    static constexpr int PTOT = C::CAP + MAXH;
    static constexpr int PER = (C::CAP + C::THREADS - 1) / C::THREADS;
    static constexpr int HL = 4 * C::HE + 4 * C::HC;  // halo entries per tile
    static constexpr size_t smem_bytes =
        sizeof(double) * (2 * PTOT + 2 * C::CAP) + sizeof(int) * (C::CAP + NC + 4) + sizeof(unsigned short) * (3 * PTOT + 8);
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
    int ntx, nty;   // tiles per side (global)
    int tr_base;    // global tile row of local row 0
    int lrow0;      // first local tile row of this launch
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

// which list of which neighbour feeds my apron: k = 0..7
//   k: 0 N-neighbour's S list | 1 S-neighbour's N list | 2 W-neighbour's E list | 3 E-neighbour's W list
//      4 NW-neighbour's SE corner | 5 NE's SW | 6 SW's NE | 7 SE's NW
// list ids: 0 N, 1 S, 2 W, 3 E, 4 NW, 5 NE, 6 SW, 7 SE
__device__ __forceinline__ void halo_source(int k, int& dr, int& dc, int& list) {
    const int drs[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
    const int dcs[8] = {0, 0, -1, 1, -1, 1, -1, 1};
    const int lists[8] = {1, 0, 3, 2, 7, 6, 5, 4};
    dr = drs[k];
    dc = dcs[k];
    list = lists[k];
}

// ------------------------------------------------------------------------------------------
// the per-step kernel
// ------------------------------------------------------------------------------------------
template <int TS, bool kStoreAcc>
__global__ void __launch_bounds__(TileCfg<TS>::THREADS) tile_step_kernel(const TileParams P) {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    constexpr int T = C::THREADS, CAP = C::CAP, W = D::W, NC = D::NC, PTOT = D::PTOT, PER = D::PER;
    constexpr int HE = C::HE, HC = C::HC, CO = C::CO;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* px = reinterpret_cast<double*>(smem_raw);
    double* py = px + PTOT;
    double* pvx = py + PTOT;
    double* pvy = pvx + CAP;
    int* pid = reinterpret_cast<int*>(pvy + CAP);
    int* ccnt = pid + CAP;  // NC + 1 (+3 pad)
    unsigned short* pcell = reinterpret_cast<unsigned short*>(ccnt + NC + 4);
    unsigned short* pslot = pcell + PTOT;
    unsigned short* sidx = pslot + PTOT;

    __shared__ int s_cnt[17];   // 0..7 halo list counts, 8..16 outbox counts of the 3x3 tiles
    __shared__ int s_hoff[9];
    __shared__ int s_warp[33];
    __shared__ int s_nown, s_nhalo, s_flags;
    __shared__ int s_hout[8];

    const int tid = threadIdx.x;
    const int lr = P.lrow0 + blockIdx.x / P.ntx, tc = blockIdx.x % P.ntx;
    const int tr = P.tr_base + lr;
    const int lt = lr * P.ntx + tc;
    const int r0 = tr * TS, c0 = tc * TS;
    const size_t gbase = (size_t)lt * CAP;
    const int n_own0 = min(P.tcount[lt], CAP);

    // ---- A0: list counts --------------------------------------------------------------------
    if (tid < 32) {
        int cnt = 0;
        if (tid < 17) {
            int dr, dc, list;
            if (tid < 8) {
                halo_source(tid, dr, dc, list);
            } else {
                dr = (tid - 8) / 3 - 1;
                dc = (tid - 8) % 3 - 1;
                list = 8;
            }
            const int ntr = tr + dr, ntc = tc + dc;
            if (ntr >= 0 && ntr < P.nty && ntc >= 0 && ntc < P.ntx) {
                const int* ec = reinterpret_cast<const int*>(row_ptr(P.exp_in, P.L, lr + dr) + P.L.off_cnt) + (size_t)ntc * 16;
                cnt = ec[list];
                cnt = min(cnt, list == 8 ? CO : halo_cap(list, HE, HC));
            }
            s_cnt[tid] = cnt;
        }
        int inc = tid < 8 ? cnt : 0;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += t;
        }
        if (tid < 8) s_hoff[tid] = inc - cnt;
        if (tid == 7) s_hoff[8] = inc;
        if (tid < 8) s_hout[tid] = 0;
        if (tid == 0) s_flags = 0;
    }
    // own particles (coalesced)
    for (int i = tid; i < n_own0; i += T) {
        px[i] = P.sx[gbase + i];
        py[i] = P.sy[gbase + i];
        pvx[i] = P.svx[gbase + i];
        pvy[i] = P.svy[gbase + i];
        pid[i] = P.sid[gbase + i];
    }
    __syncthreads();

    // ---- A1: apron from the neighbours' halo lists ---------------------------------------------
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const int cnt = s_cnt[k];
        if (cnt == 0) continue;
        int dr, dc, list;
        halo_source(k, dr, dc, list);
        const double2* src = reinterpret_cast<const double2*>(row_ptr(P.exp_in, P.L, lr + dr) + P.L.off_hxy) +
                             (size_t)(tc + dc) * D::HL + halo_offset(list, HE, HC);
        for (int e = tid; e < cnt; e += T) {
            const double2 q = src[e];
            const int h = PTOT - 1 - (s_hoff[k] + e);
            px[h] = q.x;
            py[h] = q.y;
        }
    }
    // ---- A2: outboxes of the 3x3 tiles (last warp): newcomers -> own, apron dwellers -> halo -------
    if (tid >= T - 32) {
        const int lane = tid & 31;
        const unsigned lt_mask = (1u << lane) - 1u;
        int n_own = n_own0, n_halo = s_hoff[8], flags = 0;
#pragma unroll 1
        for (int nb = 0; nb < 9; ++nb) {
            const int cnt = s_cnt[8 + nb];
            if (cnt == 0) continue;
            const int dr = nb / 3 - 1, dc = nb % 3 - 1;
            const char* row = row_ptr(P.exp_in, P.L, lr + dr);
            const size_t ob = (size_t)(tc + dc) * CO;
            const double* ox = reinterpret_cast<const double*>(row + P.L.off_ox) + ob;
            const double* oy = reinterpret_cast<const double*>(row + P.L.off_oy) + ob;
            const double* ovx = reinterpret_cast<const double*>(row + P.L.off_ovx) + ob;
            const double* ovy = reinterpret_cast<const double*>(row + P.L.off_ovy) + ob;
            const int* oid = reinterpret_cast<const int*>(row + P.L.off_oid) + ob;
            for (int e0 = 0; e0 < cnt; e0 += 32) {
                const int e = e0 + lane;
                bool mine = false, apron = false;
                double x = 0, y = 0;
                if (e < cnt) {
                    x = ox[e];
                    y = oy[e];
                    const int row_c = axis_cell(x, P.bincnt), col_c = axis_cell(y, P.bincnt);
                    mine = row_c / TS == tr && col_c / TS == tc;
                    apron = !mine && row_c >= r0 - 1 && row_c <= r0 + TS && col_c >= c0 - 1 && col_c <= c0 + TS;
                }
                const unsigned mm = __ballot_sync(0xffffffffu, mine), am = __ballot_sync(0xffffffffu, apron);
                if (mine) {
                    const int d = n_own + __popc(mm & lt_mask);
                    if (d < CAP) {
                        px[d] = x;
                        py[d] = y;
                        pvx[d] = ovx[e];
                        pvy[d] = ovy[e];
                        pid[d] = oid[e];
                    }
                }
                if (apron) {
                    const int hh = n_halo + __popc(am & lt_mask);
                    if (hh < D::MAXH) {
                        px[PTOT - 1 - hh] = x;
                        py[PTOT - 1 - hh] = y;
                    }
                }
                n_own += __popc(mm);
                n_halo += __popc(am);
            }
        }
        if (n_own > CAP) { flags |= kErrTileOverflow; n_own = CAP; }
        if (n_halo > D::MAXH) { flags |= kErrSmemOverflow; n_halo = D::MAXH; }
        if (lane == 0) {
            s_nown = n_own;
            s_nhalo = n_halo;
            if (flags) atomicOr(&s_flags, flags);
        }
    }
    for (int c = tid; c < NC + 1; c += T) ccnt[c] = 0;
    __syncthreads();
    const int n_own = s_nown, n_halo = s_nhalo, n_all = n_own + n_halo;

    // ---- B: bin own + apron particles into (TS+2)^2 cells in shared memory -----------------------
    for (int q = tid; q < n_all; q += T) {
        const int i = q < n_own ? q : PTOT - 1 - (q - n_own);
        const int lrow = axis_cell(px[i], P.bincnt) - r0 + 1, lcol = axis_cell(py[i], P.bincnt) - c0 + 1;
        if (lrow >= 0 && lrow < W && lcol >= 0 && lcol < W) {
            const int cell = lrow * W + lcol;
            pcell[i] = (unsigned short)cell;
            pslot[i] = (unsigned short)atomicAdd(&ccnt[cell], 1);
        } else {
            pcell[i] = 0xFFFFu;  // cannot happen for well-formed exports; keeps the sort safe
        }
    }
    __syncthreads();
    {
        constexpr int CHUNK = (NC + T - 1) / T;
        const int b = tid * CHUNK;
        int sum = 0;
#pragma unroll
        for (int k = 0; k < CHUNK; ++k)
            if (b + k < NC) sum += ccnt[b + k];
        int total;
        int run = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
        for (int k = 0; k < CHUNK; ++k)
            if (b + k < NC) {
                const int v = ccnt[b + k];
                ccnt[b + k] = run;
                run += v;
            }
        if (tid == 0) ccnt[NC] = total;
    }
    __syncthreads();
    for (int q = tid; q < n_all; q += T) {
        const int i = q < n_own ? q : PTOT - 1 - (q - n_own);
        const unsigned cell = pcell[i];
        if (cell != 0xFFFFu) sidx[ccnt[cell] + pslot[i]] = (unsigned short)i;
    }
    __syncthreads();

    // ---- C + D: force over the 3x3 neighbourhood, then move, all in registers ---------------------
    double nx[PER], ny[PER], nvx[PER], nvy[PER], nax[PER], nay[PER];
    int ncellrow[PER], ncellcol[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const int i = r * T + tid;
        nx[r] = ny[r] = nvx[r] = nvy[r] = nax[r] = nay[r] = 0.0;
        ncellrow[r] = ncellcol[r] = -1;
        if (i < n_own) {
            const double xi = px[i], yi = py[i];
            const int cell = pcell[i];
            auto visit = [&](auto&& f, bool want_rank) {
#pragma unroll
                for (int dr = -1; dr <= 1; ++dr) {
                    const int b = cell + dr * W;
                    const int k0 = ccnt[b - 1], k1 = ccnt[b + 2];
                    for (int k = k0; k < k1; ++k) {
                        const int j = sidx[k];
                        int rank = 0;
                        if (want_rank) rank = visit_rank(dr, (int)pcell[j] - b);
                        f(px[j], py[j], rank);
                    }
                }
            };
            double ax, ay;
            int nbc;
            accumulate_force(xi, yi, visit, ax, ay, nbc);
            double x = xi, y = yi, vx = pvx[i], vy = pvy[i];
            move_particle(x, y, vx, vy, ax, ay, P.size);
            nx[r] = x; ny[r] = y; nvx[r] = vx; nvy[r] = vy; nax[r] = ax; nay[r] = ay;
            ncellrow[r] = axis_cell(x, P.bincnt);
            ncellcol[r] = axis_cell(y, P.bincnt);
        }
    }

    // ---- E: re-tile: compact stayers in place, leavers to the outbox, boundary cells to halo lists --
    char* orow = row_ptr(P.exp_out, P.L, lr);
    double2* ohxy = reinterpret_cast<double2*>(orow + P.L.off_hxy) + (size_t)tc * D::HL;
    const size_t ob = (size_t)tc * CO;
    double* oox = reinterpret_cast<double*>(orow + P.L.off_ox) + ob;
    double* ooy = reinterpret_cast<double*>(orow + P.L.off_oy) + ob;
    double* oovx = reinterpret_cast<double*>(orow + P.L.off_ovx) + ob;
    double* oovy = reinterpret_cast<double*>(orow + P.L.off_ovy) + ob;
    double* ooax = reinterpret_cast<double*>(orow + P.L.off_oax) + ob;
    double* ooay = reinterpret_cast<double*>(orow + P.L.off_oay) + ob;
    int* ooid = reinterpret_cast<int*>(orow + P.L.off_oid) + ob;

    int stay_base = 0, leave_base = 0, flags = 0;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const int i = r * T + tid;
        const bool active = i < n_own;
        const int er = ncellrow[r] - r0, ec = ncellcol[r] - c0;
        const bool stay = active && er >= 0 && er < TS && ec >= 0 && ec < TS;
        const int code = active ? (stay ? 1 : (1 << 16)) : 0;
        int total;
        const int pre = block_exclusive_scan(code, s_warp, total);  // also orders smem reads above vs writes below
        if (active) {
            const int myid = pid[i];
            if (stay) {
                const int d = stay_base + (pre & 0xFFFF);
                P.sx[gbase + d] = nx[r];
                P.sy[gbase + d] = ny[r];
                P.svx[gbase + d] = nvx[r];
                P.svy[gbase + d] = nvy[r];
                if (kStoreAcc) {
                    P.sax[gbase + d] = nax[r];
                    P.say[gbase + d] = nay[r];
                }
                // the id stream only changes where compaction or ingestion moved a particle
                if (d >= n_own0 || d != i) P.sid[gbase + d] = myid;
                const bool n_ = er == 0, s_ = er == TS - 1, w_ = ec == 0, e_ = ec == TS - 1;
                if (n_ | s_ | w_ | e_) {
                    const double2 q = make_double2(nx[r], ny[r]);
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
            } else {
                const int d = leave_base + (pre >> 16);
                if (d < CO) {
                    oox[d] = nx[r];
                    ooy[d] = ny[r];
                    oovx[d] = nvx[r];
                    oovy[d] = nvy[r];
                    if (kStoreAcc) {
                        ooax[d] = nax[r];
                        ooay[d] = nay[r];
                    }
                    ooid[d] = myid;
                }
                const int dtr = (ncellrow[r] / TS) - tr, dtc = (ncellcol[r] / TS) - tc;
                if (dtr < -1 || dtr > 1 || dtc < -1 || dtc > 1) flags |= kErrLostParticle;
            }
        }
        stay_base += total & 0xFFFF;
        leave_base += total >> 16;
    }
    if (flags) atomicOr(&s_flags, flags);
    __syncthreads();
    if (tid < 9) {
        int* ec = reinterpret_cast<int*>(orow + P.L.off_cnt) + (size_t)tc * 16;
        if (tid < 8) {
            const int c = s_hout[tid], cap = halo_cap(tid, HE, HC);
            if (c > cap) atomicOr(&s_flags, kErrHaloOverflow);
            ec[tid] = min(c, cap);
        } else {
            if (leave_base > CO) atomicOr(&s_flags, kErrOutboxOverflow);
            ec[8] = min(leave_base, CO);
            P.tcount[lt] = stay_base;
        }
    }
    __syncthreads();
    if (tid == 0 && s_flags) atomicOr(P.err, s_flags);
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
__global__ void __launch_bounds__(TileCfg<TS>::THREADS) tile_export_kernel(const TileParams P) {
    using C = TileCfg<TS>;
    using D = TileDims<TS>;
    constexpr int T = C::THREADS, CAP = C::CAP, HE = C::HE, HC = C::HC;
    __shared__ int s_hout[8];
    __shared__ int s_flags;
    const int tid = threadIdx.x;
    const int lr = P.lrow0 + blockIdx.x / P.ntx, tc = blockIdx.x % P.ntx;
    const int tr = P.tr_base + lr;
    const int lt = lr * P.ntx + tc;
    const int r0 = tr * TS, c0 = tc * TS;
    const size_t gbase = (size_t)lt * CAP;
    const int n = min(P.tcount[lt], CAP);
    if (tid < 8) s_hout[tid] = 0;
    if (tid == 0) s_flags = 0;
    __syncthreads();
    char* orow = row_ptr(P.exp_out, P.L, lr);
    double2* ohxy = reinterpret_cast<double2*>(orow + P.L.off_hxy) + (size_t)tc * D::HL;
    for (int i = tid; i < n; i += T) {
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
        int* ec = reinterpret_cast<int*>(orow + P.L.off_cnt) + (size_t)tc * 16;
        if (tid < 8) {
            const int c = s_hout[tid], cap = halo_cap(tid, HE, HC);
            if (c > cap) atomicOr(P.err, kErrHaloOverflow);
            ec[tid] = min(c, cap);
        } else {
            ec[8] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// observation: gather the owned particles (tile stripes + outbox entries that land in owned rows)
// into a compact SoA.  One CTA per tile (ghost rows included: their outboxes may hold particles that
// have just crossed into this slab).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tile_gather_kernel(const TileParams P, int ts, int cap, int co, int lrows_alloc,
                                                          int tr_begin, int tr_end, bool have_acc, double* __restrict__ gx,
                                                          double* __restrict__ gy, double* __restrict__ gvx,
                                                          double* __restrict__ gvy, double* __restrict__ gax,
                                                          double* __restrict__ gay, int* __restrict__ gid,
                                                          int* __restrict__ cursor) {
    __shared__ int s_base, s_take[64], s_ntake;
    const int lr = blockIdx.x / P.ntx, tc = blockIdx.x % P.ntx;
    const int tr = P.tr_base + lr;
    const int lt = lr * P.ntx + tc;
    const bool owned = tr >= tr_begin && tr < tr_end;
    const int n_tile = owned ? min(P.tcount[lt], cap) : 0;
    const char* row = row_ptr(P.exp_in, P.L, lr);
    const bool row_valid = tr >= 0 && tr < P.nty;
    const int n_out = row_valid ? min((reinterpret_cast<const int*>(row + P.L.off_cnt) + (size_t)tc * 16)[8], co) : 0;
    const size_t ob = (size_t)tc * co;
    const double* ox = reinterpret_cast<const double*>(row + P.L.off_ox) + ob;
    const double* oy = reinterpret_cast<const double*>(row + P.L.off_oy) + ob;
    if (threadIdx.x == 0) {
        int k = 0;
        for (int e = 0; e < n_out && k < 64; ++e) {
            const int dtr = axis_cell(ox[e], P.bincnt) / ts;
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
    const double* ovx = reinterpret_cast<const double*>(row + P.L.off_ovx) + ob;
    const double* ovy = reinterpret_cast<const double*>(row + P.L.off_ovy) + ob;
    const double* oax = reinterpret_cast<const double*>(row + P.L.off_oax) + ob;
    const double* oay = reinterpret_cast<const double*>(row + P.L.off_oay) + ob;
    const int* oid = reinterpret_cast<const int*>(row + P.L.off_oid) + ob;
    for (int k = threadIdx.x; k < s_ntake; k += blockDim.x) {
        const int e = s_take[k], d = base + n_tile + k;
        gx[d] = ox[e];
        gy[d] = oy[e];
        gvx[d] = ovx[e];
        gvy[d] = ovy[e];
        gax[d] = have_acc ? oax[e] : 0.0;
        gay[d] = have_acc ? oay[e] : 0.0;
        gid[d] = oid[e];
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct TiledEngine {
    DeviceArena mem;
    int ts = 0, cap = 0, he = 0, hc = 0, co = 0, hl = 0, threads = 0;
    size_t smem = 0;
    int ntx = 0;             // tiles per side
    int tr_begin = 0, tr_end = 0;  // owned tile rows (global)
    int lrows = 0;           // owned rows
    int lrows_alloc = 0;     // owned + 2 ghost rows
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
    // graph replay of a parity pair
    cudaGraphExec_t graph2 = nullptr;
    bool use_graph = false;
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
    L.off_ox = take((size_t)ntx * co * sizeof(double));
    L.off_oy = take((size_t)ntx * co * sizeof(double));
    L.off_ovx = take((size_t)ntx * co * sizeof(double));
    L.off_ovy = take((size_t)ntx * co * sizeof(double));
    L.off_oax = take((size_t)ntx * co * sizeof(double));
    L.off_oay = take((size_t)ntx * co * sizeof(double));
    L.off_oid = take((size_t)ntx * co * sizeof(int));
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
    const int grid = nrows * e->ntx;
    if (store_acc)
        tile_step_kernel<TS, true><<<grid, TileCfg<TS>::THREADS, TileDims<TS>::smem_bytes, s>>>(P);
    else
        tile_step_kernel<TS, false><<<grid, TileCfg<TS>::THREADS, TileDims<TS>::smem_bytes, s>>>(P);
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
    e->threads = TileCfg<TS>::THREADS;
    e->smem = TileDims<TS>::smem_bytes;
    PSIM_CUDA(cudaFuncSetAttribute(tile_step_kernel<TS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    PSIM_CUDA(cudaFuncSetAttribute(tile_step_kernel<TS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
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
        // enough tiles to fill 148 SMs several times over, else fall to smaller tiles
        const long long cells = (long long)sim->bincnt * sim->bincnt;
        ts = cells >= 64ll * 64 * 148 * 24 ? 64 : (cells >= 32ll * 32 * 148 * 4 ? 32 : 16);
    }
    if (ts != 16 && ts != 32 && ts != 64) return fail(PSIM_ERR_INVALID, "tile_cells must be 16, 32 or 64 (got %d)", ts);
    auto* e = new TiledEngine();
    sim->tiled = e;
    e->ts = ts;
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
    if (ts == 16) tile_export_kernel<16><<<grid, TileCfg<16>::THREADS, 0, s>>>(P);
    if (ts == 32) tile_export_kernel<32><<<grid, TileCfg<32>::THREADS, 0, s>>>(P);
    if (ts == 64) tile_export_kernel<64><<<grid, TileCfg<64>::THREADS, 0, s>>>(P);
    ++sim->launches;
    PSIM_CUDA(cudaGetLastError());
    e->parity = 0;
    e->acc_valid = true;  // zeros
    e->use_graph = cfg->use_graph != 0;
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
    TileParams P = make_params(sim, e, e->parity);
    tile_gather_kernel<<<e->lrows_alloc * e->ntx, 128, 0, s>>>(P, e->ts, e->cap, e->co, e->lrows_alloc, e->tr_begin,
                                                               e->tr_end, e->acc_valid, e->g.x, e->g.y, e->g.vx, e->g.vy,
                                                               e->g.ax, e->g.ay, e->g.id, e->g_cursor);
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
    if (e->graph2) cudaGraphExecDestroy(e->graph2);
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
}

// accessors for psim_comm.cu
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
