// psim_mpi_shim.cpp -- the reference's MULTI-PROCESS flavour of the plugin interface (part2/common.h:16-32), one process per
// GPU, exported with the reference's own C++ linkage on top of the slab engine of libpsim:
//     void init_simulation  (particle_t*, int num_parts, double size, int rank, int num_procs)   _Z15init_simulationP10particle_tidii
//     void simulate_one_step(particle_t*, int num_parts, double size, int rank, int num_procs)   _Z17simulate_one_stepP10particle_tidii
//     void gather_for_save  (particle_t*, int num_parts, double size, int rank, int num_procs)   _Z15gather_for_saveP10particle_tidii
// where particle_t is part2's 56-byte record {uint64_t id; double x, y, vx, vy, ax, ay} (part2/common.h:17-25), ids 1-based.
//
// Semantics mirrored from the reference (part2/main.cpp:141-166, part2/mpi.cpp):
//   * every rank is handed the WHOLE particle array (the driver broadcasts it, main.cpp:150) and keeps the particles of its
//     1-D slab of cell rows (mpi.cpp:258-270 -> psim_slab_rows / psim_create with rank, nranks);
//   * simulate_one_step advances the slab; halo rows and migrants travel between neighbouring ranks (mpi.cpp:296-365 ->
//     GPU to GPU, fused into the step kernel over NVLink peer memory, or NCCL send/recv); the caller's array is NOT kept up to
//     date between saves (the reference does not either);
//   * gather_for_save leaves a complete, id-ordered view of all particles in rank 0's array (mpi.cpp:371-402 -> psim_gather).
// The reference needs MPI only as launcher and wire; here the ranks may be started by ANY launcher (mpirun, srun, torchrun, a
// shell loop): the only inter-process plumbing is NCCL, whose 128-byte unique id rank 0 publishes through a file
// (PSIM_RENDEZVOUS=<path>, default /tmp/psim_rendezvous_<parent pid>) that the other ranks poll.  No MPI symbols are needed.
//
// Steps are counted and enqueued in batches (PSIM_FLUSH calls, default 30; always before a gather and on call nsteps-1) so
// that the kstep engine can fuse them -- psim_step(n) is bit-identical to n calls of psim_step(1).
//
// Environment: PSIM_DEVICE (default LOCAL_RANK, else rank % visible devices), PSIM_ENGINE=kstep|tiled, PSIM_TILE, PSIM_FLUSH,
//              PSIM_RENDEZVOUS, PSIM_VERBOSE.
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace capi {   // the C ABI (its 48-byte particle_t stays in this namespace; C linkage ignores it)
#include "../../include/psim.h"
}

// reference part2/common.h:17-25
typedef struct particle_t {
    uint64_t id;
    double x, y, vx, vy, ax, ay;
} particle_t;

namespace {

capi::psim_sim* g_sim = nullptr;
long long g_call = 0;
int g_pending = 0, g_flush_every = 30;
std::vector<capi::particle_t> g_stage;   // 48-byte view of the caller's array (create / gather staging)

[[noreturn]] void die(int status, const char* file, int line) {
    std::fprintf(stderr, "GPUassert: %s: %s %s %d\n", capi::psim_error_string(status), capi::psim_last_error(), file, line);
    std::exit(status ? status : 1);
}
#define SHIM_CHECK(call)                         \
    do {                                         \
        int st__ = (call);                       \
        if (st__ != PSIM_OK) die(st__, __FILE__, __LINE__); \
    } while (0)

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

std::string rendezvous_path() {
    if (const char* p = std::getenv("PSIM_RENDEZVOUS")) return p;
    return "/tmp/psim_rendezvous_" + std::to_string((long long)getppid());
}

// rank 0 writes the id (temp file + rename: readers never see a partial file), the others poll for it
void exchange_unique_id(unsigned char id[128], int rank) {
    const std::string path = rendezvous_path();
    if (rank == 0) {
        SHIM_CHECK(capi::psim_comm_unique_id(id));
        const std::string tmp = path + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f || std::fwrite(id, 1, 128, f) != 128 || std::fclose(f) != 0 || std::rename(tmp.c_str(), path.c_str()) != 0) {
            std::fprintf(stderr, "psim_mpi_shim: cannot publish the NCCL id in %s\n", path.c_str());
            std::exit(1);
        }
        return;
    }
    for (int spin = 0; spin < 60000; ++spin) {   // up to 60 s
        if (FILE* f = std::fopen(path.c_str(), "rb")) {
            const size_t got = std::fread(id, 1, 128, f);
            std::fclose(f);
            if (got == 128) return;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    std::fprintf(stderr, "psim_mpi_shim: rank %d found no NCCL id in %s (set PSIM_RENDEZVOUS to a path all ranks share)\n", rank, path.c_str());
    std::exit(1);
}

void flush(bool store_acc) {
    if (g_pending > 0) SHIM_CHECK(capi::psim_step(g_sim, g_pending, store_acc ? PSIM_STEP_DEFAULT : PSIM_STEP_ACCEL_NONE));
    g_pending = 0;
}

}  // namespace

void init_simulation(particle_t* parts, int num_parts, double size, int rank, int num_procs) {
    capi::psim_config cfg;
    capi::psim_config_default(&cfg);
    cfg.engine = PSIM_ENGINE_KSTEP;
    if (const char* e = std::getenv("PSIM_ENGINE"))
        if (!std::strcmp(e, "tiled")) cfg.engine = PSIM_ENGINE_TILED;
    cfg.tile_cells = env_int("PSIM_TILE", 0);
    cfg.rank = rank;
    cfg.nranks = num_procs;
    cfg.device = env_int("PSIM_DEVICE", env_int("LOCAL_RANK", -2));
    if (cfg.device == -2) {   // no launcher hint: spread the ranks over the visible devices
        int ndev = 1;
        if (capi::psim_device_count(&ndev) != PSIM_OK || ndev < 1) ndev = 1;
        cfg.device = rank % ndev;
    }
    g_flush_every = env_int("PSIM_FLUSH", 30);
    if (g_flush_every < 1) g_flush_every = 1;
    if (g_sim) {
        capi::psim_destroy(g_sim);
        g_sim = nullptr;
    }
    g_call = 0;
    g_pending = 0;
    // the engine identifies a particle by its position in the array; part2's id (1-based, main.cpp: init_particles) must agree
    g_stage.resize((size_t)num_parts);
    for (int i = 0; i < num_parts; ++i) {
        if (parts[i].id != (uint64_t)i + 1) {
            std::fprintf(stderr, "psim_mpi_shim: parts[%d].id == %llu, expected %d (ids must be 1..N in array order)\n", i,
                         (unsigned long long)parts[i].id, i + 1);
            std::exit(1);
        }
        g_stage[(size_t)i] = capi::particle_t{parts[i].x, parts[i].y, parts[i].vx, parts[i].vy, 0.0, 0.0};
    }
    SHIM_CHECK(capi::psim_create(&g_sim, &cfg, g_stage.data(), num_parts, size));
    if (num_procs > 1) {
        unsigned char id[128];
        exchange_unique_id(id, rank);
        SHIM_CHECK(capi::psim_comm_connect(g_sim, id));   // collective: when it returns everybody has read the file
        if (rank == 0) std::remove(rendezvous_path().c_str());
    }
    if (rank != 0) std::vector<capi::particle_t>().swap(g_stage);   // only the root gathers
    if (env_int("PSIM_VERBOSE", 0)) {
        capi::psim_info_t info;
        capi::psim_info(g_sim, &info);
        std::fprintf(stderr, "[psim mpi-flavour] rank %d/%d device %d rows %d..%d tile %d steps/launch %d\n", rank, num_procs, info.device,
                     info.row_begin, info.row_end, info.tile_cells, info.steps_per_launch);
    }
}

void simulate_one_step(particle_t* /*parts*/, int /*num_parts*/, double /*size*/, int /*rank*/, int /*num_procs*/) {
    if (!g_sim) die(PSIM_ERR_STATE, __FILE__, __LINE__);
    const long long k = g_call++;
    ++g_pending;
    if (g_pending >= g_flush_every || k == PSIM_NSTEPS - 1) {
        flush(false);
        SHIM_CHECK(capi::psim_sync(g_sim));   // surfaces device-side errors; also ends the driver's timed region honestly
    }
}

void gather_for_save(particle_t* parts, int num_parts, double /*size*/, int rank, int /*num_procs*/) {
    if (!g_sim) die(PSIM_ERR_STATE, __FILE__, __LINE__);
    flush(true);
    SHIM_CHECK(capi::psim_gather(g_sim, rank == 0 ? g_stage.data() : nullptr, 0));
    if (rank == 0)
        for (int i = 0; i < num_parts; ++i) {
            const capi::particle_t& q = g_stage[(size_t)i];
            parts[i].x = q.x; parts[i].y = q.y; parts[i].vx = q.vx; parts[i].vy = q.vy; parts[i].ax = q.ax; parts[i].ay = q.ay;
        }
}
