// psim_comm.cpp -- slab exchange over NCCL (one process per GPU).
//
// A slab's halo exchange and particle migration are one message per neighbour per step: the byte
// range that holds the exports (halo lists + outboxes, psim_tiled.h) of its first / last owned tile
// row goes to the neighbour's ghost row (precedent: the four MPI_Sendrecv of reference
// part2/mpi.cpp:136-140,241-245 -- ghost rows before the forces, migrants after the move -- fused
// here into a single grouped send/recv because migrants travel inside the exports).
//
// NCCL is bound at run time (dlopen) so that libpsim has no link-time dependency on a particular
// NCCL build: inside a torch process `libnccl.so.2` resolves to the copy torch already loaded.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <vector>

#include "psim_internal.h"

namespace psim {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

// Communicators are process-level plumbing (the analogue of MPI_COMM_WORLD, which the reference's MPI driver sets up
// before its timer, part2/main.cpp): a communicator created for a given unique id and slab geometry is kept for the life
// of the process and shared by every handle that connects with the same id, so that a second simulation does not pay the
// NCCL bootstrap (about a second) again.
struct CachedComm {
    std::string key;
    ncclComm_t comm;
};
static std::vector<CachedComm> g_comms;
static std::mutex g_comms_mutex;

static int load_nccl() {
    if (g_nccl.lib) return PSIM_OK;
    const char* names[] = {std::getenv("PSIM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(PSIM_ERR_COMM, "cannot load NCCL (set PSIM_NCCL_LIB): %s", dlerror());
#define PSIM_SYM(field, name)                                                            \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(lib, name);                         \
    if (!g_nccl.field) return fail(PSIM_ERR_COMM, "NCCL symbol %s missing", name);
    PSIM_SYM(GetUniqueId, "ncclGetUniqueId")
    PSIM_SYM(CommInitRank, "ncclCommInitRank")
    PSIM_SYM(CommDestroy, "ncclCommDestroy")
    PSIM_SYM(Send, "ncclSend")
    PSIM_SYM(Recv, "ncclRecv")
    PSIM_SYM(GroupStart, "ncclGroupStart")
    PSIM_SYM(GroupEnd, "ncclGroupEnd")
    PSIM_SYM(AllReduce, "ncclAllReduce")
    PSIM_SYM(GetErrorString, "ncclGetErrorString")
#undef PSIM_SYM
    g_nccl.lib = lib;
    return PSIM_OK;
}

#define PSIM_NCCL(call)                                                                                   \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) return fail(PSIM_ERR_COMM, "%s: %s", #call, g_nccl.GetErrorString(r__));  \
    } while (0)

int tiled_exchange(psim_sim* sim, int parity, cudaStream_t s) {
    if (sim->nranks == 1) return PSIM_OK;
    static const bool skip = std::getenv("PSIM_DEBUG_SKIP_EXCHANGE") != nullptr;   // timing experiments only: results are wrong
    if (skip) return PSIM_OK;
    if (!sim->comm) return fail(PSIM_ERR_STATE, "slab %d/%d is not connected: call psim_comm_connect first", sim->rank, sim->nranks);
    ncclComm_t comm = static_cast<ncclComm_t>(sim->comm);
    char *first, *last, *glo, *ghi;
    size_t bytes;
    tiled_boundary_rows(sim, parity, &first, &last, &glo, &ghi, &bytes);
    PSIM_NCCL(g_nccl.GroupStart());
    if (sim->rank > 0) {
        PSIM_NCCL(g_nccl.Send(first, bytes, ncclInt8, sim->rank - 1, comm, s));
        PSIM_NCCL(g_nccl.Recv(glo, bytes, ncclInt8, sim->rank - 1, comm, s));
    }
    if (sim->rank < sim->nranks - 1) {
        PSIM_NCCL(g_nccl.Send(last, bytes, ncclInt8, sim->rank + 1, comm, s));
        PSIM_NCCL(g_nccl.Recv(ghi, bytes, ncclInt8, sim->rank + 1, comm, s));
    }
    PSIM_NCCL(g_nccl.GroupEnd());
    return PSIM_OK;
}

// kstep engine: a tile row is four byte ranges (headers, pos, vel, id stripes); the first / last owned row goes to the
// neighbour's ghost row.  Used for the first exchange after create and, per launch, by the NCCL flavour.
int kstep_exchange(psim_sim* sim, int parity, cudaStream_t s) {
    if (sim->nranks == 1) return PSIM_OK;
    if (!sim->comm) return fail(PSIM_ERR_STATE, "slab %d/%d is not connected: call psim_comm_connect first", sim->rank, sim->nranks);
    ncclComm_t comm = static_cast<ncclComm_t>(sim->comm);
    const int lrows = kstep_owned_rows(sim);
    char *first[4], *last[4], *glo[4], *ghi[4];
    size_t bytes[4];
    kstep_row_ranges(sim, parity, 1, first, bytes);
    kstep_row_ranges(sim, parity, lrows, last, bytes);
    kstep_row_ranges(sim, parity, 0, glo, bytes);
    kstep_row_ranges(sim, parity, lrows + 1, ghi, bytes);
    PSIM_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < 4; ++k) {
        if (sim->rank > 0) {
            PSIM_NCCL(g_nccl.Send(first[k], bytes[k], ncclInt8, sim->rank - 1, comm, s));
            PSIM_NCCL(g_nccl.Recv(glo[k], bytes[k], ncclInt8, sim->rank - 1, comm, s));
        }
        if (sim->rank < sim->nranks - 1) {
            PSIM_NCCL(g_nccl.Send(last[k], bytes[k], ncclInt8, sim->rank + 1, comm, s));
            PSIM_NCCL(g_nccl.Recv(ghi[k], bytes[k], ncclInt8, sim->rank + 1, comm, s));
        }
    }
    PSIM_NCCL(g_nccl.GroupEnd());
    return PSIM_OK;
}

// ---- gather to one rank (the reference's gather_for_save, part2/mpi.cpp:371-402) ------------------------------------------
// counts[r] = number of records rank r owns (sum-allreduce of a vector in which every rank fills only its own slot)
int comm_allgather_counts(psim_sim* sim, int mine, int* counts_host, cudaStream_t s) {
    for (int r = 0; r < sim->nranks; ++r) counts_host[r] = r == sim->rank ? mine : 0;
    if (sim->nranks == 1) return PSIM_OK;
    if (!sim->comm) return fail(PSIM_ERR_STATE, "slab %d/%d is not connected: call psim_comm_connect first", sim->rank, sim->nranks);
    ncclComm_t comm = static_cast<ncclComm_t>(sim->comm);
    int* d = nullptr;
    PSIM_CUDA(cudaMalloc(&d, sizeof(int) * (size_t)sim->nranks));
    cudaError_t e = cudaMemcpyAsync(d, counts_host, sizeof(int) * (size_t)sim->nranks, cudaMemcpyHostToDevice, s);
    ncclResult_t r = e == cudaSuccess ? g_nccl.AllReduce(d, d, (size_t)sim->nranks, ncclInt32, ncclSum, comm, s) : ncclSuccess;
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(counts_host, d, sizeof(int) * (size_t)sim->nranks, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (r != ncclSuccess) return fail(PSIM_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    if (e != cudaSuccess) return fail(PSIM_ERR_CUDA, "comm_allgather_counts: %s", cudaGetErrorString(e));
    return PSIM_OK;
}

// sum-allreduce of a small host vector (in place)
int comm_allreduce_ints(psim_sim* sim, int* host_values, int count, cudaStream_t s) {
    if (sim->nranks == 1) return PSIM_OK;
    if (!sim->comm) return fail(PSIM_ERR_STATE, "slab %d/%d is not connected", sim->rank, sim->nranks);
    ncclComm_t comm = static_cast<ncclComm_t>(sim->comm);
    int* d = nullptr;
    PSIM_CUDA(cudaMalloc(&d, sizeof(int) * (size_t)count));
    cudaError_t e = cudaMemcpyAsync(d, host_values, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, s);
    ncclResult_t r = e == cudaSuccess ? g_nccl.AllReduce(d, d, (size_t)count, ncclInt32, ncclSum, comm, s) : ncclSuccess;
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(host_values, d, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (r != ncclSuccess) return fail(PSIM_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    if (e != cudaSuccess) return fail(PSIM_ERR_CUDA, "comm_allreduce_ints: %s", cudaGetErrorString(e));
    return PSIM_OK;
}

// all-to-all of variable-size record blocks (+ their ids): I send block [send_offs[r], send_offs[r+1]) to rank r and receive
// count_matrix[src * R + me] records from every src, packed in rank order.  The block for myself is a device copy.
int comm_alltoall_records(psim_sim* sim, const void* send, const int* send_ids, const int* send_offs, void* recv, int* recv_ids,
                          const int* count_matrix, size_t rec_bytes, cudaStream_t s) {
    const int R = sim->nranks, me = sim->rank;
    ncclComm_t comm = static_cast<ncclComm_t>(sim->comm);
    size_t roff = 0;
    PSIM_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < R; ++r) {
        const size_t n_in = (size_t)count_matrix[(size_t)r * R + me], n_out = (size_t)(send_offs[r + 1] - send_offs[r]);
        if (r != me) {
            if (n_out) {
                PSIM_NCCL(g_nccl.Send(static_cast<const char*>(send) + (size_t)send_offs[r] * rec_bytes, n_out * rec_bytes, ncclInt8, r, comm, s));
                PSIM_NCCL(g_nccl.Send(send_ids + send_offs[r], n_out * sizeof(int), ncclInt8, r, comm, s));
            }
            if (n_in) {
                PSIM_NCCL(g_nccl.Recv(static_cast<char*>(recv) + roff * rec_bytes, n_in * rec_bytes, ncclInt8, r, comm, s));
                PSIM_NCCL(g_nccl.Recv(recv_ids + roff, n_in * sizeof(int), ncclInt8, r, comm, s));
            }
        }
        roff += n_in;
    }
    PSIM_NCCL(g_nccl.GroupEnd());
    roff = 0;
    for (int r = 0; r < me; ++r) roff += (size_t)count_matrix[(size_t)r * R + me];
    const size_t n_self = (size_t)(send_offs[me + 1] - send_offs[me]);
    if (n_self) {
        PSIM_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + roff * rec_bytes, static_cast<const char*>(send) + (size_t)send_offs[me] * rec_bytes,
                                  n_self * rec_bytes, cudaMemcpyDeviceToDevice, s));
        PSIM_CUDA(cudaMemcpyAsync(recv_ids + roff, send_ids + send_offs[me], n_self * sizeof(int), cudaMemcpyDeviceToDevice, s));
    }
    return PSIM_OK;
}

// Every rank but `root` sends its `mine` packed records (device) + ids (device); root receives rank r's block at offset
// offs[r] of rec_all / id_all (device arrays sized for all particles).  One NCCL group.
int comm_gather_records(psim_sim* sim, int root, const void* rec, const int* ids, int mine, size_t rec_bytes, void* rec_all, int* id_all,
                        const int* counts, cudaStream_t s) {
    if (sim->nranks == 1) return PSIM_OK;
    ncclComm_t comm = static_cast<ncclComm_t>(sim->comm);
    PSIM_NCCL(g_nccl.GroupStart());
    if (sim->rank != root) {
        if (mine > 0) {
            PSIM_NCCL(g_nccl.Send(rec, (size_t)mine * rec_bytes, ncclInt8, root, comm, s));
            PSIM_NCCL(g_nccl.Send(ids, (size_t)mine * sizeof(int), ncclInt8, root, comm, s));
        }
    } else {
        size_t off = 0;
        for (int r = 0; r < sim->nranks; ++r) {
            if (r != root && counts[r] > 0) {
                PSIM_NCCL(g_nccl.Recv(static_cast<char*>(rec_all) + off * rec_bytes, (size_t)counts[r] * rec_bytes, ncclInt8, r, comm, s));
                PSIM_NCCL(g_nccl.Recv(id_all + off, (size_t)counts[r] * sizeof(int), ncclInt8, r, comm, s));
            }
            off += (size_t)counts[r];
        }
    }
    PSIM_NCCL(g_nccl.GroupEnd());
    return PSIM_OK;
}

// ---- peer-memory exchange ----------------------------------------------------------------------------
// Each slab maps its neighbours' export buffers and flag words (CUDA IPC; the 64-byte handles travel through the
// NCCL communicator once at connect time).  Per step: the kernel stores its boundary rows' exports into the
// neighbours' ghost rows, a one-thread kernel then writes my step count into their flag words, and my next step
// is ordered behind THEIR flags with cuStreamWaitValue32 (no SM, no host round trip).
struct P2PHandles {
    cudaIpcMemHandle_t exports[2];
    cudaIpcMemHandle_t flags;
};

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static WaitValue32Fn g_wait_value32 = nullptr;

static int p2p_setup(psim_sim* sim, ncclComm_t comm) {
    const char* env = std::getenv("PSIM_P2P");
    if (env && env[0] == '0') return PSIM_OK;
    char *e0, *e1;
    size_t bytes, row_bytes;
    int lrows, ntx;
    if (sim->kstep) kstep_shared_buffers(sim, &e0, &e1, &bytes, &ntx);
    else tiled_export_buffers(sim, &e0, &e1, &bytes, &row_bytes, &lrows, &ntx);
    if (ntx / sim->nranks < 2) return PSIM_OK;   // (same decision on every rank) a one-row slab would have to mirror the same row to both sides: keep NCCL
    (void)lrows;
    (void)row_bytes;
    if (!g_wait_value32) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            cudaGetLastError();
            return PSIM_OK;   // no stream memory operations: keep NCCL
        }
        g_wait_value32 = reinterpret_cast<WaitValue32Fn>(fn);
    }
    PSIM_CUDA(cudaMalloc(&sim->d_flags, 2 * sizeof(int)));
    PSIM_CUDA(cudaMemset(sim->d_flags, 0, 2 * sizeof(int)));
    P2PHandles mine{}, theirs[2]{};
    PSIM_CUDA(cudaIpcGetMemHandle(&mine.exports[0], e0));
    PSIM_CUDA(cudaIpcGetMemHandle(&mine.exports[1], e1));
    PSIM_CUDA(cudaIpcGetMemHandle(&mine.flags, sim->d_flags));
    // swap handles with the neighbours (through device buffers: NCCL moves device memory)
    P2PHandles* d_buf = nullptr;   // [0] mine, [1] from below, [2] from above
    PSIM_CUDA(cudaMalloc(&d_buf, 3 * sizeof(P2PHandles)));
    PSIM_CUDA(cudaMemcpy(d_buf, &mine, sizeof mine, cudaMemcpyHostToDevice));
    cudaStream_t s = sim->comm_stream;
    PSIM_NCCL(g_nccl.GroupStart());
    if (sim->rank > 0) {
        PSIM_NCCL(g_nccl.Send(d_buf, sizeof(P2PHandles), ncclInt8, sim->rank - 1, comm, s));
        PSIM_NCCL(g_nccl.Recv(d_buf + 1, sizeof(P2PHandles), ncclInt8, sim->rank - 1, comm, s));
    }
    if (sim->rank < sim->nranks - 1) {
        PSIM_NCCL(g_nccl.Send(d_buf, sizeof(P2PHandles), ncclInt8, sim->rank + 1, comm, s));
        PSIM_NCCL(g_nccl.Recv(d_buf + 2, sizeof(P2PHandles), ncclInt8, sim->rank + 1, comm, s));
    }
    PSIM_NCCL(g_nccl.GroupEnd());
    PSIM_CUDA(cudaStreamSynchronize(s));
    PSIM_CUDA(cudaMemcpy(theirs, d_buf + 1, 2 * sizeof(P2PHandles), cudaMemcpyDeviceToHost));
    cudaFree(d_buf);
    // map the neighbours' buffers; the outcome is agreed on by ALL slabs (min over ranks), so that either every rank uses the
    // peer-memory exchange or every rank stays on NCCL send / recv -- a one-sided failure must not leave a neighbour waiting
    int ok = 1;
    for (int side = 0; side < 2 && ok; ++side) {
        const bool have = side == 0 ? sim->rank > 0 : sim->rank < sim->nranks - 1;
        if (!have) continue;
        for (int b = 0; b < 2 && ok; ++b) {
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, theirs[side].exports[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
            else sim->peer_exports[side][b] = static_cast<char*>(p);
        }
        void* f = nullptr;
        if (ok && cudaIpcOpenMemHandle(&f, theirs[side].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
        else if (ok) sim->peer_flags[side] = static_cast<int*>(f);
    }
    if (!ok) cudaGetLastError();
    {
        int* d_ok = nullptr;
        PSIM_CUDA(cudaMalloc(&d_ok, sizeof(int)));
        PSIM_CUDA(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
        PSIM_NCCL(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, comm, s));
        PSIM_CUDA(cudaStreamSynchronize(s));
        PSIM_CUDA(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
        cudaFree(d_ok);
    }
    if (!ok) {   // somebody could not map a neighbour: everybody keeps the NCCL flavour
        for (int side = 0; side < 2; ++side) {
            for (int b = 0; b < 2; ++b) {
                if (sim->peer_exports[side][b]) cudaIpcCloseMemHandle(sim->peer_exports[side][b]);
                sim->peer_exports[side][b] = nullptr;
            }
            if (sim->peer_flags[side]) cudaIpcCloseMemHandle(sim->peer_flags[side]);
            sim->peer_flags[side] = nullptr;
        }
        cudaFree(sim->d_flags);
        sim->d_flags = nullptr;
        return PSIM_OK;
    }
    for (int k = 0; k < 2; ++k) {
        PSIM_CUDA(cudaEventCreateWithFlags(&sim->ev_b[k], cudaEventDisableTiming));
        PSIM_CUDA(cudaEventCreateWithFlags(&sim->ev_i[k], cudaEventDisableTiming));
    }
    sim->p2p = true;
    sim->p2p_steps = 0;
    return PSIM_OK;
}

int comm_p2p_wait(psim_sim* sim, cudaStream_t s) {
    if (!sim->p2p || sim->p2p_steps == 0) return PSIM_OK;
    for (int side = 0; side < 2; ++side) {
        const bool have = side == 0 ? sim->rank > 0 : sim->rank < sim->nranks - 1;
        if (!have) continue;
        const CUresult r = g_wait_value32(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(sim->d_flags + side),
                                          sim->p2p_steps, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) return fail(PSIM_ERR_COMM, "cuStreamWaitValue32 failed (%d)", (int)r);
    }
    return PSIM_OK;
}

int comm_p2p_signal(psim_sim* sim, cudaStream_t s) {
    if (!sim->p2p) return PSIM_OK;
    ++sim->p2p_steps;
    // my lower neighbour sees me as ITS upper neighbour (flag word 1) and vice versa
    launch_flag_store(sim->peer_flags[0] ? sim->peer_flags[0] + 1 : nullptr, sim->peer_flags[1] ? sim->peer_flags[1] + 0 : nullptr,
                      (int)sim->p2p_steps, s);
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

void comm_destroy(psim_sim* sim) {
    if (sim->p2p && sim->d_flags) {
        // the neighbours store into my ghost rows until they have finished the same step: do not unmap / free before
        // (host-side poll with a time-out: a neighbour that ran fewer steps must not hang this process for ever)
        cudaStreamSynchronize(sim->stream);
        for (int spin = 0; spin < 20000; ++spin) {
            int f[2] = {0, 0};
            if (cudaMemcpy(f, sim->d_flags, sizeof f, cudaMemcpyDeviceToHost) != cudaSuccess) break;
            const bool lo_ok = sim->rank == 0 || (unsigned)f[0] >= sim->p2p_steps;
            const bool hi_ok = sim->rank == sim->nranks - 1 || (unsigned)f[1] >= sim->p2p_steps;
            if (lo_ok && hi_ok) break;
            struct timespec ts = {0, 500000};
            nanosleep(&ts, nullptr);
        }
    }
    for (int side = 0; side < 2; ++side) {
        for (int b = 0; b < 2; ++b)
            if (sim->peer_exports[side][b]) cudaIpcCloseMemHandle(sim->peer_exports[side][b]);
        if (sim->peer_flags[side]) cudaIpcCloseMemHandle(sim->peer_flags[side]);
        sim->peer_exports[side][0] = sim->peer_exports[side][1] = nullptr;
        sim->peer_flags[side] = nullptr;
    }
    if (sim->d_flags) cudaFree(sim->d_flags);
    sim->d_flags = nullptr;
    for (int k = 0; k < 2; ++k) {
        if (sim->ev_b[k]) cudaEventDestroy(sim->ev_b[k]);
        if (sim->ev_i[k]) cudaEventDestroy(sim->ev_i[k]);
        sim->ev_b[k] = sim->ev_i[k] = nullptr;
    }
    sim->p2p = false;
    sim->comm = nullptr;   // the communicator itself stays cached for the process (see g_comms)
    if (sim->ev_boundary) cudaEventDestroy(sim->ev_boundary);
    if (sim->ev_exchanged) cudaEventDestroy(sim->ev_exchanged);
    if (sim->comm_stream) cudaStreamDestroy(sim->comm_stream);
    sim->ev_boundary = sim->ev_exchanged = nullptr;
    sim->comm_stream = nullptr;
}

}  // namespace psim

using namespace psim;

extern "C" int psim_comm_unique_id(unsigned char id128[128]) {
    if (!id128) return fail(PSIM_ERR_INVALID, "psim_comm_unique_id: NULL");
    PSIM_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    PSIM_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return PSIM_OK;
}

extern "C" int psim_comm_connect(psim_sim* sim, const unsigned char id128[128]) {
    if (!sim || !id128) return fail(PSIM_ERR_INVALID, "psim_comm_connect: NULL argument");
    if (sim->nranks == 1) return PSIM_OK;
    if (sim->comm) return fail(PSIM_ERR_STATE, "psim_comm_connect: already connected");
    PSIM_TRY(load_nccl());
    PSIM_CUDA(cudaSetDevice(sim->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_comms_mutex);
        std::string key(reinterpret_cast<const char*>(id128), 128);
        key += "/" + std::to_string(sim->rank) + "/" + std::to_string(sim->nranks) + "/" + std::to_string(sim->device);
        for (const CachedComm& c : g_comms)
            if (c.key == key) comm = c.comm;
        if (!comm) {
            PSIM_NCCL(g_nccl.CommInitRank(&comm, sim->nranks, id, sim->rank));
            g_comms.push_back({key, comm});
        }
    }
    sim->comm = comm;
    int lo = 0, hi = 0;
    PSIM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PSIM_CUDA(cudaStreamCreateWithPriority(&sim->comm_stream, cudaStreamNonBlocking, hi));   // exchange first
    PSIM_CUDA(cudaEventCreateWithFlags(&sim->ev_boundary, cudaEventDisableTiming));
    PSIM_CUDA(cudaEventCreateWithFlags(&sim->ev_exchanged, cudaEventDisableTiming));
    static const bool trace = std::getenv("PSIM_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    if (kstep_fill_pending(sim)) PSIM_TRY(kstep_distributed_fill(sim));   // cooperative upload: needs the communicator
    const double t1 = now();
    if (sim->tiled || sim->kstep) PSIM_TRY(p2p_setup(sim, comm));
    if (trace)
        std::fprintf(stderr, "[psim trace] connect rank %d: cooperative fill %.3f ms, peer-memory setup %.3f ms\n", sim->rank, 1e3 * (t1 - t0),
                     1e3 * (now() - t1));
    return PSIM_OK;
}
