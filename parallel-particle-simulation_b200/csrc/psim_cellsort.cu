// psim_cellsort.cu -- the counting-sort binner and the "cellsort" engine.
//
// Per time step (reference part3/gpu.cu:187-208 does reset/rebin/forces/move with fixed 16-slot
// bins; this is a different design, not a port):
//   1. hist    : cell = floor(x/0.01)*B + floor(y/0.01); slot = atomicAdd(count[cell], 1)
//   2. scan    : single-pass decoupled-look-back exclusive scan of the counts (warp shuffles)
//   3. scatter : particle -> cell_start[cell] + slot in a cell-sorted structure of arrays
//   4. force+move : one thread per particle walks its 3x3 neighbourhood (three contiguous index
//                   ranges, one per cell row), accumulates, integrates, bounces, writes the new
//                   state to the other SoA buffer.
// No per-cell capacity exists anywhere, so any particle distribution is handled.
#include "psim_force.cuh"
#include "psim_internal.h"

namespace psim {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 256;

// (1) atomic histogram over cutoff cells; remembers the arrival rank inside the cell.
// Two particles per thread through 16-byte loads when the pair is in range.
__global__ void __launch_bounds__(kThreads) hist_cells_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                              int n, int bincnt, int* __restrict__ count,
                                                              int* __restrict__ slot) {
    const int i0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i0 + 1 < n) {
        const double2 xv = *reinterpret_cast<const double2*>(x + i0);
        const double2 yv = *reinterpret_cast<const double2*>(y + i0);
        const int c0 = axis_cell(xv.x, bincnt) * bincnt + axis_cell(yv.x, bincnt);
        const int c1 = axis_cell(xv.y, bincnt) * bincnt + axis_cell(yv.y, bincnt);
        int2 s;
        s.x = atomicAdd(count + c0, 1);
        s.y = atomicAdd(count + c1, 1);
        *reinterpret_cast<int2*>(slot + i0) = s;
    } else if (i0 < n) {
        const int c0 = axis_cell(x[i0], bincnt) * bincnt + axis_cell(y[i0], bincnt);
        slot[i0] = atomicAdd(count + c0, 1);
    }
}

// (2) exclusive scan, in place, single pass with decoupled look-back.
constexpr int kScanItems = 16;
constexpr int kScanTile = kThreads * kScanItems;
constexpr unsigned long long kFlagAgg = 1ull << 32, kFlagPrefix = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads) scan_cells_kernel(int* __restrict__ data, long long n_items,
                                                              unsigned long long* __restrict__ desc,
                                                              int* __restrict__ ticket) {
    __shared__ int s_tile, s_prefix;
    __shared__ int s_warp[33];
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
    int v[kScanItems];
    if (base + kScanItems <= n_items) {
#pragma unroll
        for (int k = 0; k < kScanItems; k += 4) {
            const int4 q = *reinterpret_cast<const int4*>(data + base + k);
            v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < n_items) ? data[base + k] : 0;
    }
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum += v[k];
    int total;
    int tprefix = block_exclusive_scan(tsum, s_warp, total);

    if (threadIdx.x == 0) {
        st_desc(desc + tile, (tile == 0 ? kFlagPrefix : kFlagAgg) | (unsigned)total);
        if (tile == 0) s_prefix = 0;
    }
    if (tile > 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int exclusive = 0;
        for (int look = tile - 1;; look -= 32) {
            const int idx = look - lane;
            unsigned long long d = idx >= 0 ? ld_desc(desc + idx) : kFlagPrefix;
            while (__any_sync(0xffffffffu, (d >> 32) == 0)) {
                if ((d >> 32) == 0) d = ld_desc(desc + idx);
            }
            const unsigned pm = __ballot_sync(0xffffffffu, (d >> 32) == 2);
            const int first = pm ? __ffs(pm) - 1 : 32;
            int contrib = lane <= first ? (int)(unsigned)(d & 0xffffffffu) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            exclusive += contrib;
            if (pm) break;
        }
        if (lane == 0) {
            st_desc(desc + tile, kFlagPrefix | (unsigned)(exclusive + total));
            s_prefix = exclusive;
        }
    }
    __syncthreads();
    int run = s_prefix + tprefix;
    if (base + kScanItems <= n_items) {
#pragma unroll
        for (int k = 0; k < kScanItems; k += 4) {
            int4 q;
            q.x = run; run += v[k];
            q.y = run; run += v[k + 1];
            q.z = run; run += v[k + 2];
            q.w = run; run += v[k + 3];
            *reinterpret_cast<int4*>(data + base + k) = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (base + k < n_items) data[base + k] = run;
            run += v[k];
        }
    }
}

// (3) scatter into the cell-sorted SoA.  The cell is recomputed (two IEEE divisions) instead of
// being stored and re-read: FP64 issue slots are cheaper than 8 bytes of HBM traffic here.
__global__ void __launch_bounds__(kThreads) scatter_cells_kernel(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ vx,
    const double* __restrict__ vy, const int* __restrict__ id, const int* __restrict__ slot, int n, int bincnt,
    const int* __restrict__ cell_start, double* __restrict__ ox, double* __restrict__ oy, double* __restrict__ ovx,
    double* __restrict__ ovy, int* __restrict__ oid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = x[i], yi = y[i];
    const int c = axis_cell(xi, bincnt) * bincnt + axis_cell(yi, bincnt);
    const int d = cell_start[c] + slot[i];
    ox[d] = xi;
    oy[d] = yi;
    ovx[d] = vx[i];
    ovy[d] = vy[i];
    oid[d] = id[i];
}

// (4) force + move over the cell-sorted SoA (in: sorted buffer, out: the other buffer, same index).
template <bool kStoreAcc>
__global__ void __launch_bounds__(kThreads) force_move_cells_kernel(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ vx,
    const double* __restrict__ vy, int n, int bincnt, double size, const int* __restrict__ cell_start,
    double* __restrict__ ox, double* __restrict__ oy, double* __restrict__ ovx, double* __restrict__ ovy,
    double* __restrict__ oax, double* __restrict__ oay) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xi = x[i], yi = y[i];
    const int row = axis_cell(xi, bincnt), col = axis_cell(yi, bincnt);
    const int c_lo = max(col - 1, 0), c_hi = min(col + 1, bincnt - 1);
    int k0[3], k1[3];
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr) {
        const int rr = row + dr;
        if (rr < 0 || rr >= bincnt) {
            k0[dr + 1] = k1[dr + 1] = 0;
        } else {
            const long long b = (long long)rr * bincnt;
            k0[dr + 1] = __ldg(cell_start + b + c_lo);
            k1[dr + 1] = __ldg(cell_start + b + c_hi + 1);
        }
    }
    auto visit = [&](auto&& f) {
#pragma unroll 1
        for (int dr = -1; dr <= 1; ++dr) {
            for (int k = k0[dr + 1]; k < k1[dr + 1]; ++k) {
                const double xj = __ldg(x + k), yj = __ldg(y + k);
                f(xj, yj, dr);
            }
        }
    };
    auto rank_of = [&](double, double yj, int dr) { return visit_rank(dr, axis_cell(yj, bincnt) - col); };
    double ax, ay;
    int nb;
    accumulate_force(xi, yi, visit, rank_of, ax, ay, nb);
    double vxi = vx[i], vyi = vy[i];
    move_particle(xi, yi, vxi, vyi, ax, ay, size);
    ox[i] = xi;
    oy[i] = yi;
    ovx[i] = vxi;
    ovy[i] = vyi;
    if (kStoreAcc) {
        oax[i] = ax;
        oay[i] = ay;
    }
}

// ------------------------------------------------------------------------------------------
// CellBinner
// ------------------------------------------------------------------------------------------
int CellBinner::init(int bincnt_, int capacity) {
    bincnt = bincnt_;
    ncell = (long long)bincnt * bincnt;
    capacity_n = capacity;
    scan_tiles = (int)((ncell + 1 + kScanTile - 1) / kScanTile);
    PSIM_TRY(mem.alloc(&cell_start, (size_t)ncell + 1));
    PSIM_TRY(mem.alloc(&slot, (size_t)capacity + 2));
    PSIM_TRY(mem.alloc(&scan_desc, (size_t)scan_tiles));
    PSIM_TRY(mem.alloc(&scan_ticket, 1));
    return PSIM_OK;
}

int CellBinner::count(const double* x, const double* y, int n, cudaStream_t s) {
    if (n > capacity_n) return fail(PSIM_ERR_INVALID, "CellBinner: n=%d exceeds capacity %d", n, capacity_n);
    PSIM_CUDA(cudaMemsetAsync(cell_start, 0, sizeof(int) * ((size_t)ncell + 1), s));
    if (n > 0) {
        const int blocks = ((n + 1) / 2 + kThreads - 1) / kThreads;
        hist_cells_kernel<<<blocks, kThreads, 0, s>>>(x, y, n, bincnt, cell_start, slot);
        ++launches;
    }
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

int CellBinner::build(const double* x, const double* y, int n, cudaStream_t s) {
    PSIM_TRY(count(x, y, n, s));
    PSIM_CUDA(cudaMemsetAsync(scan_desc, 0, sizeof(unsigned long long) * (size_t)scan_tiles, s));
    PSIM_CUDA(cudaMemsetAsync(scan_ticket, 0, sizeof(int), s));
    scan_cells_kernel<<<scan_tiles, kThreads, 0, s>>>(cell_start, ncell + 1, scan_desc, scan_ticket);
    ++launches;
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

// ------------------------------------------------------------------------------------------
// CellsortEngine
// ------------------------------------------------------------------------------------------
struct CellsortEngine {
    DeviceArena mem;
    CellBinner binner;
    int n = 0;
    // cur: state after the last step (ordered by the previous step's cells); srt: cell-sorted scratch
    double *x[2] = {}, *y[2] = {}, *vx[2] = {}, *vy[2] = {};
    int* id[2] = {};
    double *ax = nullptr, *ay = nullptr;  // acceleration of the last stored step, indexed like buffer `cur`
    int cur = 0;
    bool acc_valid = true;   // ax / ay belong to the step that produced the current order (false after a step that did not store them)
};

__global__ void __launch_bounds__(kThreads) aos_to_soa_kernel(const particle_t* __restrict__ p, int n,
                                                              double* __restrict__ x, double* __restrict__ y,
                                                              double* __restrict__ vx, double* __restrict__ vy,
                                                              double* __restrict__ ax, double* __restrict__ ay,
                                                              int* __restrict__ id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2* q = reinterpret_cast<const double2*>(p + i);
    const double2 a = q[0], b = q[1];
    x[i] = a.x;
    y[i] = a.y;
    vx[i] = b.x;
    vy[i] = b.y;
    ax[i] = 0.0;
    ay[i] = 0.0;
    id[i] = i;
}

int cellsort_create(psim_sim* sim, const particle_t* d_parts, int n) {
    auto* e = new CellsortEngine();
    sim->cs = e;
    e->n = n;
    PSIM_TRY(e->binner.init(sim->bincnt, n));
    for (int b = 0; b < 2; ++b) {
        PSIM_TRY(e->mem.alloc(&e->x[b], (size_t)n + 2));
        PSIM_TRY(e->mem.alloc(&e->y[b], (size_t)n + 2));
        PSIM_TRY(e->mem.alloc(&e->vx[b], (size_t)n + 2));
        PSIM_TRY(e->mem.alloc(&e->vy[b], (size_t)n + 2));
        PSIM_TRY(e->mem.alloc(&e->id[b], (size_t)n + 2));
    }
    PSIM_TRY(e->mem.alloc(&e->ax, (size_t)n + 2));
    PSIM_TRY(e->mem.alloc(&e->ay, (size_t)n + 2));
    if (n > 0) {
        aos_to_soa_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(
            d_parts, n, e->x[0], e->y[0], e->vx[0], e->vy[0], e->ax, e->ay, e->id[0]);
        ++sim->launches;
        PSIM_CUDA(cudaGetLastError());
    }
    return PSIM_OK;
}

int cellsort_step(psim_sim* sim, int nsteps, int flags) {
    CellsortEngine* e = sim->cs;
    const int n = e->n;
    cudaStream_t s = sim->stream;
    const int blocks = (n + kThreads - 1) / kThreads;
    for (int step = 0; step < nsteps; ++step) {
        const bool store = (flags & PSIM_STEP_ACCEL_ALL) || (!(flags & PSIM_STEP_ACCEL_NONE) && step == nsteps - 1);
        const int a = e->cur, b = a ^ 1;
        const long long before = e->binner.launches;
        PSIM_TRY(e->binner.build(e->x[a], e->y[a], n, s));
        sim->launches += e->binner.launches - before;
        if (n > 0) {
            scatter_cells_kernel<<<blocks, kThreads, 0, s>>>(e->x[a], e->y[a], e->vx[a], e->vy[a], e->id[a],
                                                             e->binner.slot, n, sim->bincnt, e->binner.cell_start,
                                                             e->x[b], e->y[b], e->vx[b], e->vy[b], e->id[b]);
            if (store)
                force_move_cells_kernel<true><<<blocks, kThreads, 0, s>>>(
                    e->x[b], e->y[b], e->vx[b], e->vy[b], n, sim->bincnt, sim->size, e->binner.cell_start, e->x[a],
                    e->y[a], e->vx[a], e->vy[a], e->ax, e->ay);
            else
                force_move_cells_kernel<false><<<blocks, kThreads, 0, s>>>(
                    e->x[b], e->y[b], e->vx[b], e->vy[b], n, sim->bincnt, sim->size, e->binner.cell_start, e->x[a],
                    e->y[a], e->vx[a], e->vy[a], e->ax, e->ay);
            sim->launches += 2;
            // ids follow the sorted order: the new state in buffer a is indexed like buffer b
            std::swap(e->id[a], e->id[b]);
        }
        PSIM_CUDA(cudaGetLastError());
        e->acc_valid = store;   // a step that does not store leaves ax / ay indexed by an older order: stale
        ++sim->steps_done;
    }
    return PSIM_OK;
}

bool cellsort_acc_valid(psim_sim* sim) { return sim->cs && sim->cs->acc_valid; }

int cellsort_view(psim_sim* sim, SoAView* out) {
    CellsortEngine* e = sim->cs;
    const int a = e->cur;
    out->x = e->x[a];
    out->y = e->y[a];
    out->vx = e->vx[a];
    out->vy = e->vy[a];
    out->ax = e->ax;
    out->ay = e->ay;
    out->id = e->id[a];
    out->n = e->n;
    return PSIM_OK;
}

void cellsort_destroy(psim_sim* sim) {
    if (!sim->cs) return;
    sim->cs->binner.release();
    sim->cs->mem.release();
    delete sim->cs;
    sim->cs = nullptr;
}

long long cellsort_bytes(psim_sim* sim) { return sim->cs ? (long long)(sim->cs->mem.bytes + sim->cs->binner.mem.bytes) : 0; }

}  // namespace psim
