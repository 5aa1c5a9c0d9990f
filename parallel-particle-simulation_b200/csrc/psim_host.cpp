// psim_host.cpp -- host-side driver helpers behind the C ABI: the particle generator and the
// trajectory writer.  Both must reproduce the reference driver's observable behaviour exactly
// (generator: same engine, same draw order, same distributions => same bits as reference
// part1/main.cpp:31-59 when built against libstdc++; writer: the text format of part1/main.cpp:15-28).
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "psim_internal.h"

extern "C" int psim_init_particles(particle_t* parts, int num_parts, double size, int seed) {
    if (num_parts < 0 || (num_parts > 0 && !parts)) return psim::fail(PSIM_ERR_INVALID, "psim_init_particles: bad array");
    std::random_device entropy;
    std::mt19937 rng(seed ? (unsigned)seed : entropy());

    // lattice of sx columns by sy rows; every particle takes one free lattice site at random
    const int sx = (int)std::ceil(std::sqrt((double)num_parts));
    const int sy = sx ? (num_parts + sx - 1) / sx : 0;
    std::vector<int> free_site((size_t)num_parts);
    for (int k = 0; k < num_parts; ++k) free_site[(size_t)k] = k;

    for (int i = 0, remaining = num_parts; i < num_parts; ++i, --remaining) {
        std::uniform_int_distribution<int> pick(0, remaining - 1);
        const int j = pick(rng);
        const int site = free_site[(size_t)j];
        free_site[(size_t)j] = free_site[(size_t)remaining - 1];

        particle_t& p = parts[i];
        p.x = size * (1. + (site % sx)) / (1 + sx);
        p.y = size * (1. + (site / sx)) / (1 + sy);
        std::uniform_real_distribution<float> speed(-1.0, 1.0);
        p.vx = speed(rng);
        p.vy = speed(rng);
        p.ax = 0.0;
        p.ay = 0.0;
    }
    return PSIM_OK;
}

extern "C" int psim_save_frame(void* file, const double* xy, int num_parts, double size, int first) {
    if (!file || (num_parts > 0 && !xy)) return psim::fail(PSIM_ERR_INVALID, "psim_save_frame: NULL argument");
    FILE* f = static_cast<FILE*>(file);
    // default ostream formatting of the reference (precision 6, neither fixed nor scientific) == %g
    if (first) std::fprintf(f, "%d %g\n", num_parts, size);
    for (int i = 0; i < num_parts; ++i) std::fprintf(f, "%g %g\n", xy[2 * i], xy[2 * i + 1]);
    std::fputc('\n', f);
    return std::ferror(f) ? psim::fail(PSIM_ERR_INVALID, "psim_save_frame: write failed") : PSIM_OK;
}
