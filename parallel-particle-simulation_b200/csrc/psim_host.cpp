// psim_host.cpp -- host-side driver helpers behind the C ABI: the particle generator and the
// trajectory writer.  Both must reproduce the reference driver's observable behaviour exactly
// (generator: same engine, same draw order, same distributions => same bits as reference
// part1/main.cpp:31-59 when built against libstdc++; writer: the text format of part1/main.cpp:15-28).
#include <algorithm>
#include <charconv>
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

// One double in the reference's stream formatting (operator<< with the default precision 6 and neither fixed nor
// scientific == printf("%g")).  std::to_chars with chars_format::general and precision 6 is specified to produce
// exactly the characters of printf("%.6g") in the C locale, several times faster than the stdio formatter.
static inline char* put_g6(char* p, char* end, double v) {
    const std::to_chars_result r = std::to_chars(p, end, v, std::chars_format::general, 6);
    return r.ptr;
}

extern "C" int psim_save_frame(void* file, const double* xy, int num_parts, double size, int first) {
    if (!file || (num_parts > 0 && !xy)) return psim::fail(PSIM_ERR_INVALID, "psim_save_frame: NULL argument");
    FILE* f = static_cast<FILE*>(file);
    if (first) std::fprintf(f, "%d %g\n", num_parts, size);
    // format into a large buffer, one fwrite per chunk (the reference flushes every line through std::endl,
    // part1/main.cpp:23: same bytes, far more system calls)
    constexpr size_t kChunk = 1 << 16;                 // particles per buffer
    std::vector<char> buf(kChunk * 64 + 64);           // two numbers of <= 24 characters, a blank and a newline
    for (size_t i0 = 0; i0 < (size_t)num_parts; i0 += kChunk) {
        const size_t i1 = std::min(i0 + kChunk, (size_t)num_parts);
        char* p = buf.data();
        char* const end = p + buf.size();
        for (size_t i = i0; i < i1; ++i) {
            p = put_g6(p, end, xy[2 * i]);
            *p++ = ' ';
            p = put_g6(p, end, xy[2 * i + 1]);
            *p++ = '\n';
        }
        if (std::fwrite(buf.data(), 1, (size_t)(p - buf.data()), f) != (size_t)(p - buf.data()))
            return psim::fail(PSIM_ERR_INVALID, "psim_save_frame: write failed");
    }
    std::fputc('\n', f);
    return std::ferror(f) ? psim::fail(PSIM_ERR_INVALID, "psim_save_frame: write failed") : PSIM_OK;
}
