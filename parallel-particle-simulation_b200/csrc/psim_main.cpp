// psim_main.cpp -- this repository's own driver: the reference driver's command line, stdout line
// and trajectory format (reference part1/main.cpp:95-150 / part3/main.cu:96-151), driving the
// CUDA engines through the C ABI.
//
//   psim [-h] [-n <particles>=1000] [-s <seed>=0] [-o <trajectory file>]
//
// Timed region, as in the reference (main.cpp:120-144): init (psim_create, which includes the
// host->device upload) + nsteps steps + the saves, ended by a device synchronisation.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <iostream>
#include <vector>

#include "../../include/psim.h"

static int find_flag(int argc, char** argv, const char* flag) {
    for (int i = 1; i < argc; ++i)
        if (!std::strcmp(argv[i], flag)) return i;
    return -1;
}
static int int_option(int argc, char** argv, const char* flag, int dflt) {
    const int at = find_flag(argc, argv, flag);
    return (at >= 0 && at + 1 < argc) ? std::atoi(argv[at + 1]) : dflt;
}
static const char* str_option(int argc, char** argv, const char* flag) {
    const int at = find_flag(argc, argv, flag);
    return (at >= 0 && at + 1 < argc) ? argv[at + 1] : nullptr;
}

#define CHECK(call)                                                                                        \
    do {                                                                                                   \
        int st__ = (call);                                                                                 \
        if (st__ != PSIM_OK) {                                                                             \
            std::fprintf(stderr, "GPUassert: %s: %s %s %d\n", psim_error_string(st__), psim_last_error(), __FILE__, __LINE__); \
            return st__;                                                                                   \
        }                                                                                                  \
    } while (0)

int main(int argc, char** argv) {
    if (find_flag(argc, argv, "-h") >= 0) {
        std::cout << "Options:\n-h: see this help\n-n <int>: set number of particles\n"
                     "-o <filename>: set the output file name\n-s <int>: set particle initialization seed\n";
        return 0;
    }
    const char* savename = str_option(argc, argv, "-o");
    FILE* fsave = savename ? std::fopen(savename, "w") : nullptr;
    if (fsave) std::setvbuf(fsave, nullptr, _IOFBF, 1 << 22);

    const int num_parts = int_option(argc, argv, "-n", 1000);
    const int seed = int_option(argc, argv, "-s", 0);
    const double size = std::sqrt(PSIM_DENSITY * num_parts);

    std::vector<particle_t> parts((size_t)num_parts);
    CHECK(psim_init_particles(parts.data(), num_parts, size, seed));

    psim_config cfg;
    psim_config_default(&cfg);
    if (const char* e = std::getenv("PSIM_ENGINE")) {
        if (!std::strcmp(e, "cellsort")) cfg.engine = PSIM_ENGINE_CELLSORT;
        else if (!std::strcmp(e, "tiled")) cfg.engine = PSIM_ENGINE_TILED;
        else if (!std::strcmp(e, "kstep")) cfg.engine = PSIM_ENGINE_KSTEP;
    }
    if (const char* t = std::getenv("PSIM_TILE")) cfg.tile_cells = std::atoi(t);

    CHECK(psim_device_init(-1));   // context before the clock, like the reference's CUDA driver (part3/main.cu:120-125)
    const auto t0 = std::chrono::steady_clock::now();
    psim_sim* sim = nullptr;
    CHECK(psim_create(&sim, &cfg, parts.data(), num_parts, size));
    // The steps between two saves are enqueued as one batch (the kstep engine fuses them; psim_step(n) is bit-identical to n
    // calls of psim_step(1)).  Saves are pipelined three ways: frame k is gathered on the device and copied to a page-locked
    // host buffer asynchronously (psim_read_positions_begin) while the GPU already runs the next batch; a writer thread formats
    // frame k (the dominant cost of -o at large N) while frame k+1 is on its way into the other buffer.
    std::vector<double> xy[2];
    if (fsave)
        for (int b = 0; b < 2; ++b) {
            xy[b].resize((size_t)num_parts * 2);
            (void)psim_host_register(xy[b].data(), xy[b].size() * sizeof(double));   // best effort
        }
    std::future<int> writing[2];   // the writer that last used buffer b
    int cur = 0;
    bool have_frame = false, first = true;
    int frame_buf = 0;
    auto finish_frame = [&]() -> int {   // frame in flight: wait for its copy, hand it to the writer (frames stay in order)
        if (!have_frame) return PSIM_OK;
        int st = psim_read_positions_end(sim);
        if (st != PSIM_OK) return st;
        const int other = frame_buf ^ 1;
        if (writing[other].valid() && (st = writing[other].get()) != PSIM_OK) return st;
        writing[frame_buf] = std::async(std::launch::async, psim_save_frame, (void*)fsave, (const double*)xy[frame_buf].data(), num_parts, size,
                                        first ? 1 : 0);
        first = false;
        have_frame = false;
        return PSIM_OK;
    };
    for (int step = 0; step < PSIM_NSTEPS;) {
        // next step after which the state is observed: the save steps 0, 10, 20, ... (reference main.cpp:135), else the end
        int stop = PSIM_NSTEPS - 1;
        if (fsave) stop = std::min(stop, step % PSIM_SAVEFREQ == 0 ? step : (step / PSIM_SAVEFREQ + 1) * PSIM_SAVEFREQ);
        const int batch = stop - step + 1;
        CHECK(psim_step(sim, batch, stop == PSIM_NSTEPS - 1 ? PSIM_STEP_DEFAULT : PSIM_STEP_ACCEL_NONE));
        step += batch;
        if (fsave && (stop % PSIM_SAVEFREQ) == 0) {
            if (writing[cur].valid()) CHECK(writing[cur].get());   // the buffer is free again
            CHECK(psim_read_positions_begin(sim, xy[cur].data()));   // asynchronous: the next batch is enqueued behind it right away
            CHECK(finish_frame());                                    // (the PREVIOUS frame, if any)
            have_frame = true;
            frame_buf = cur;
            cur ^= 1;
        }
    }
    CHECK(finish_frame());
    for (int b = 0; b < 2; ++b)
        if (writing[b].valid()) CHECK(writing[b].get());
    CHECK(psim_sync(sim));
    const auto t1 = std::chrono::steady_clock::now();
    const double seconds = std::chrono::duration<double>(t1 - t0).count();

    std::cout << "Simulation Time = " << seconds << " seconds for " << num_parts << " particles.\n";
    if (std::getenv("PSIM_VERBOSE")) {
        psim_info_t info;
        psim_info(sim, &info);
        std::fprintf(stderr, "[psim] engine=%s tile=%d launches=%lld particle-steps/s=%.4g\n",
                     info.engine == PSIM_ENGINE_KSTEP ? "kstep" : info.engine == PSIM_ENGINE_TILED ? "tiled" : "cellsort", info.tile_cells, info.kernel_launches,
                     (double)num_parts * PSIM_NSTEPS / seconds);
    }
    if (fsave) std::fclose(fsave);
    psim_destroy(sim);
    return 0;
}
