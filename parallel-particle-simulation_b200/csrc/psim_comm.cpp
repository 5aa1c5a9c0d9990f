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
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

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
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

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

void comm_destroy(psim_sim* sim) {
    if (sim->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(sim->comm));
    sim->comm = nullptr;
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
    PSIM_NCCL(g_nccl.CommInitRank(&comm, sim->nranks, id, sim->rank));
    sim->comm = comm;
    int lo = 0, hi = 0;
    PSIM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PSIM_CUDA(cudaStreamCreateWithPriority(&sim->comm_stream, cudaStreamNonBlocking, hi));   // exchange first
    PSIM_CUDA(cudaEventCreateWithFlags(&sim->ev_boundary, cudaEventDisableTiming));
    PSIM_CUDA(cudaEventCreateWithFlags(&sim->ev_exchanged, cudaEventDisableTiming));
    return PSIM_OK;
}
