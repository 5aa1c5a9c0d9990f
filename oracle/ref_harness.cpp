// ref_harness.cpp -- K-step timing / state-dump harness around the UNMODIFIED reference
// kernels (TEST / BASELINE INFRASTRUCTURE, not product code).
//
// The reference's stock drivers hard-code 1000 steps (`#define nsteps 1000`,
// reference part1/common.h:5), which at 20 M particles is ~25 min of OpenMP work.
// This harness links the reference's own part1/openmp.cpp (or serial.cpp) and its
// own particle generator (init_particles from part1/main.cpp, compiled with
// -Dmain=ref_stock_main so that it can be linked as a library), and runs W warm-up
// plus K timed steps exactly the way part1/main.cpp:124-139 does: every thread of
// one `#pragma omp parallel` region calls simulate_one_step each step.
//
// Built only by oracle/Makefile, from sources where they lie under /root/reference,
// into oracle/_ref/ (git-ignored).  Output: one JSON object on stdout.
//
//   ref_harness_openmp -n N -s SEED -k STEPS [-w WARMUP] [-d dumpfile]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "common.h"  // the reference's own header (-I/root/reference/part1)

void init_particles(particle_t* parts, int num_parts, double size, int part_seed);  // part1/main.cpp:31

static long arg_long(int argc, char** argv, const char* flag, long dflt) {
    for (int i = 1; i + 1 < argc; ++i)
        if (!std::strcmp(argv[i], flag)) return std::atol(argv[i + 1]);
    return dflt;
}
static const char* arg_str(int argc, char** argv, const char* flag) {
    for (int i = 1; i + 1 < argc; ++i)
        if (!std::strcmp(argv[i], flag)) return argv[i + 1];
    return nullptr;
}

int main(int argc, char** argv) {
    const int n = (int)arg_long(argc, argv, "-n", 1000);
    const int seed = (int)arg_long(argc, argv, "-s", 1);
    const int k = (int)arg_long(argc, argv, "-k", 10);
    const int w = (int)arg_long(argc, argv, "-w", 0);
    const char* dump = arg_str(argc, argv, "-d");
    const double size = std::sqrt(density * n);  // part1/main.cpp:113

    particle_t* parts = new particle_t[n];
    init_particles(parts, n, size, seed);
    for (int i = 0; i < n; ++i) parts[i].ax = parts[i].ay = 0;  // the stock driver leaves these unset

    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    init_simulation(parts, n, size);
    auto t1 = clk::now();
    clk::time_point t2 = t1, t3 = t1;
    int threads = 1;
#ifdef _OPENMP
#pragma omp parallel default(shared)
#endif
    {
#ifdef _OPENMP
#pragma omp master
        threads = omp_get_num_threads();
#endif
        for (int step = 0; step < w + k; ++step) {
            if (step == w) {
#ifdef _OPENMP
#pragma omp barrier
#pragma omp master
#endif
                t2 = clk::now();
            }
            simulate_one_step(parts, n, size);
        }
#ifdef _OPENMP
#pragma omp barrier
#pragma omp master
#endif
        t3 = clk::now();
    }
    const double init_s = std::chrono::duration<double>(t1 - t0).count();
    const double steps_s = std::chrono::duration<double>(t3 - t2).count();
    if (dump) {
        FILE* f = std::fopen(dump, "wb");
        if (!f) { std::perror("dump"); return 2; }
        std::fwrite(parts, sizeof(particle_t), (size_t)n, f);
        std::fclose(f);
    }
    std::printf("{\"n\": %d, \"seed\": %d, \"size\": %.17g, \"threads\": %d, \"warmup\": %d, \"steps\": %d, "
                "\"init_s\": %.6f, \"steps_s\": %.6f, \"particle_steps_per_s\": %.6g}\n",
                n, seed, size, threads, w, k, init_s, steps_s, steps_s > 0 ? (double)n * k / steps_s : 0.0);
    delete[] parts;
    return 0;
}
