// psim_internal.h -- host-side structures shared by the engines and the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/psim.h"

namespace psim {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);

#define PSIM_CUDA(call)                                                                                 \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return ::psim::fail(PSIM_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                                __LINE__);                                                              \
    } while (0)

#define PSIM_TRY(expr)                 \
    do {                               \
        int s__ = (expr);              \
        if (s__ != PSIM_OK) return s__; \
    } while (0)

// device error word (sticky, OR-ed bit flags) ----------------------------------------------------
enum : int {
    kErrTileOverflow   = 1,   // a tile received more particles than it has slots
    kErrHaloOverflow   = 2,   // an edge / corner halo list overflowed
    kErrOutboxOverflow = 4,   // more particles left a tile in one step than its outbox holds
    kErrSmemOverflow   = 8,   // a tile's working set (own + apron) exceeded the shared-memory staging area
    kErrLostParticle   = 16,  // a particle moved farther than one tile in one step
    kErrSpeedBound     = 32,  // kstep engine: a velocity exceeded the bound that keeps K fused steps exact (replayed with K = 1)
};
constexpr int kErrWords = 16;   // device error block: [0] flags, [1..7] high-water marks / debug, [8] first failed launch + 1

// ---- device buffer bookkeeping ---------------------------------------------------------------
struct DeviceArena {
    std::vector<void*> ptrs;
    size_t bytes = 0;
    // optional: one big block that the following alloc() calls carve up (a dozen multi-hundred-megabyte cudaMallocs cost
    // tens of milliseconds inside the drivers' timed region; one does not)
    char* block = nullptr;
    size_t block_size = 0, block_used = 0;
    int reserve(size_t total) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, total);
        if (e != cudaSuccess) return fail(PSIM_ERR_CUDA, "cudaMalloc(%zu bytes): %s", total, cudaGetErrorString(e));
        ptrs.push_back(q);
        block = static_cast<char*>(q);
        block_size = total;
        block_used = 0;
        bytes += total;
        return PSIM_OK;
    }
    template <class T>
    int alloc(T** p, size_t count) {
        *p = nullptr;
        if (count == 0) count = 1;
        const size_t need = (count * sizeof(T) + 255) & ~(size_t)255;
        if (block && block_used + need <= block_size) {
            *p = reinterpret_cast<T*>(block + block_used);
            block_used += need;
            return PSIM_OK;
        }
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T));
        if (e != cudaSuccess) return fail(PSIM_ERR_CUDA, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        ptrs.push_back(q);
        bytes += count * sizeof(T);
        *p = static_cast<T*>(q);
        return PSIM_OK;
    }
    void release() {
        for (void* q : ptrs) cudaFree(q);
        ptrs.clear();
        bytes = 0;
        block = nullptr;
        block_size = block_used = 0;
    }
};

// Compact structure-of-arrays view of the particles a handle owns (device pointers).
struct SoAView {
    double *x = nullptr, *y = nullptr, *vx = nullptr, *vy = nullptr, *ax = nullptr, *ay = nullptr;
    int* id = nullptr;
    int n = 0;
};

// ---- cell binner: counting sort over cutoff cells (psim_cellsort.cu) ----------------------------
// histogram (atomic) -> single-pass decoupled-look-back exclusive scan -> scatter.
struct CellBinner {
    DeviceArena mem;
    int bincnt = 0;
    long long ncell = 0;
    int capacity_n = 0;
    int* cell_start = nullptr;   // ncell + 1 : counts, then exclusive prefix in place
    int* slot = nullptr;         // n : rank of a particle inside its cell (arrival order)
    unsigned long long* scan_desc = nullptr;
    int* scan_ticket = nullptr;
    int scan_tiles = 0;
    long long launches = 0;
    int init(int bincnt, int capacity_n);
    // per-cell counts into cell_start[0..ncell) and the arrival rank of every particle into slot
    int count(const double* x, const double* y, int n, cudaStream_t s);
    // count + in-place exclusive scan: cell_start[c] = first sorted index of cell c, cell_start[ncell] = n
    int build(const double* x, const double* y, int n, cudaStream_t s);
    void release() { mem.release(); }
};

struct CellsortEngine;
struct TiledEngine;
struct KstepEngine;

}  // namespace psim

struct psim_sim {
    int engine = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int n_total = 0;  // records in the caller's array
    double size = 0;
    int bincnt = 0;
    int rank = 0, nranks = 1;
    int row_begin = 0, row_end = 0;
    long long steps_done = 0;
    long long launches = 0;
    bool input_on_device = false;
    int owned_last = 0;       // slabs: particles owned at the last observation (view) call
    // asynchronous position read-back (psim_read_positions_begin / _end): write-back kernel on the handle's stream, D2H on a
    // copy stream, two device staging buffers so that a read can fly while the next steps run
    cudaStream_t copy_stream = nullptr;
    double2* async_stage[2] = {nullptr, nullptr};
    cudaEvent_t async_ready[2] = {nullptr, nullptr}, async_done[2] = {nullptr, nullptr};
    unsigned async_issued = 0, async_waited = 0;
    double* async_dst[2] = {nullptr, nullptr};   // host destination of the read in each slot
    int* async_err = nullptr;                    // pinned: device error words as of each read, [slot][kErrWords]
    int engine_switches = 0;  // kstep -> cellsort hand-overs after an unrecoverable capacity / speed failure
    int* d_err = nullptr;  // device error word
    int* h_err = nullptr;  // pinned mirror
    psim::DeviceArena mem;
    psim::CellsortEngine* cs = nullptr;
    psim::TiledEngine* tiled = nullptr;
    psim::KstepEngine* kstep = nullptr;
    // scratch for observation calls
    psim::DeviceArena scratch;
    char* scratch_ptr = nullptr;
    size_t scratch_bytes = 0;
    void* comm = nullptr;  // ncclComm_t when connected
    // slab exchange runs on its own stream so that it overlaps the interior tile rows of the same step
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_boundary = nullptr;   // boundary tile rows of the current step are done (compute stream)
    cudaEvent_t ev_exchanged = nullptr;  // ghost rows hold the neighbours' exports (comm stream)
    // peer-memory exchange (default for slabs on one NVSwitch box): the step kernel stores the exports of its boundary
    // tile rows straight into the neighbour's ghost rows; a flag per neighbour orders the steps
    bool p2p = false;
    int* d_flags = nullptr;                 // [0] written by the lower neighbour, [1] by the upper: its completed step count
    char* peer_exports[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side][parity]: neighbour's export buffers, mapped
    int* peer_flags[2] = {nullptr, nullptr};                               // [side]: neighbour's d_flags, mapped
    unsigned p2p_steps = 0;                 // steps signalled so far
    cudaEvent_t ev_b[2] = {nullptr, nullptr};   // boundary launch of step s done (index s & 1)
    cudaEvent_t ev_i[2] = {nullptr, nullptr};   // interior launch of step s done (index s & 1)
};

namespace psim {

// engines (each returns PSIM_* status)
int cellsort_create(psim_sim* sim, const particle_t* d_parts_aos, int n);
int cellsort_step(psim_sim* sim, int nsteps, int flags);
int cellsort_view(psim_sim* sim, SoAView* out);  // current compact state
bool cellsort_acc_valid(psim_sim* sim);            // ax / ay of the view belong to the last step
void cellsort_destroy(psim_sim* sim);
long long cellsort_bytes(psim_sim* sim);

int tiled_create(psim_sim* sim, const psim_config* cfg, const particle_t* parts, int n, bool parts_on_device,
                 bool* unsuitable);
int tiled_exchange(psim_sim* sim, int parity, cudaStream_t s);  // psim_comm.cpp
void tiled_boundary_rows(psim_sim* sim, int parity, char** first_owned, char** last_owned, char** ghost_lo, char** ghost_hi,
                         size_t* row_bytes);
void comm_destroy(psim_sim* sim);
int comm_allgather_counts(psim_sim* sim, int mine, int* counts_host, cudaStream_t s);
int comm_gather_records(psim_sim* sim, int root, const void* rec, const int* ids, int mine, size_t rec_bytes, void* rec_all, int* id_all,
                        const int* counts, cudaStream_t s);
int comm_p2p_wait(psim_sim* sim, cudaStream_t s);     // stream waits until both neighbours have finished step p2p_steps
int comm_p2p_signal(psim_sim* sim, cudaStream_t s);   // tell both neighbours that my step ++p2p_steps is finished
void tiled_export_buffers(psim_sim* sim, char** parity0, char** parity1, size_t* bytes, size_t* row_bytes, int* lrows, int* ntx);
void launch_flag_store(int* flag_a, int* flag_b, int value, cudaStream_t s);   // psim_tiled.cu
int tiled_step(psim_sim* sim, int nsteps, int flags);
int tiled_view(psim_sim* sim, SoAView* out);  // gathers into scratch
int tiled_writeback(psim_sim* sim, particle_t* d_out, double2* d_xy);  // original-order write-back straight from the tiles
void tiled_destroy(psim_sim* sim);
long long tiled_bytes(psim_sim* sim);
void tiled_info(psim_sim* sim, psim_info_t* out);
int tiled_tile_rows(int bincnt, int ts);
void tiled_slab_rows(int ntx, int rank, int nranks, int* begin, int* end);

// kstep engine (psim_kstep.cu): K steps fused per launch in shared memory
int kstep_create(psim_sim* sim, const psim_config* cfg, const particle_t* parts, int n, bool parts_on_device,
                 bool* unsuitable);
int kstep_step(psim_sim* sim, int nsteps, int flags);
int kstep_after_sync(psim_sim* sim, bool* replayed, int* pending_steps, bool* pending_store);   // error words are in sim->h_err, stream idle
int kstep_view(psim_sim* sim, SoAView* out);
int kstep_writeback(psim_sim* sim, particle_t* d_out, double2* d_xy);
void kstep_destroy(psim_sim* sim);
long long kstep_bytes(psim_sim* sim);
void kstep_info(psim_sim* sim, psim_info_t* out);
int kstep_default_tile(int bincnt);
bool kstep_tile_supported(int ts);   // 16, 32, 64 and one build-time tunable size (psim_kstep.cu)
int kstep_exchange(psim_sim* sim, int parity, cudaStream_t s);  // psim_comm.cpp
void kstep_shared_buffers(psim_sim* sim, char** parity0, char** parity1, size_t* bytes, int* ntx);
void kstep_row_ranges(psim_sim* sim, int parity, int lrow, char* ptr[4], size_t bytes[4]);
int kstep_owned_rows(psim_sim* sim);
int kstep_ghost_capacity(psim_sim* sim);
int kstep_ghost_positions(psim_sim* sim, double* gx, double* gy, int capacity, int* count);   // neighbours for slab statistics
bool kstep_fill_pending(psim_sim* sim);      // cooperative upload deferred to psim_comm_connect
int kstep_distributed_fill(psim_sim* sim);
int comm_allreduce_ints(psim_sim* sim, int* host_values, int count, cudaStream_t s);   // in place, sum
int comm_alltoall_records(psim_sim* sim, const void* send, const int* send_ids, const int* send_offs, void* recv, int* recv_ids,
                          const int* count_matrix, size_t rec_bytes, cudaStream_t s);

}  // namespace psim
