// ref_probe.cpp -- read-only C accessors into the UNMODIFIED reference kernel's global bin
// structure (TEST INFRASTRUCTURE).  Linked into oracle/_ref/libref_{serial,openmp}.so next to
// the reference's own part1/serial.cpp / openmp.cpp, which define
//     int BinCnt;  std::unordered_set<particle_t*>* Bins;      (reference part1/serial.cpp:14-16)
// so that the tests can compare per-cell counts and membership lists bit-exactly against the
// data structure the reference itself maintains, and can see the reference's in-cell visiting
// order (hash-set iteration order).
#include <unordered_set>
#include "common.h"  // the reference's header

extern int BinCnt;
extern std::unordered_set<particle_t*>* Bins;

extern "C" {
int ref_bin_count() { return BinCnt; }

// counts[c] = size of the reference's set for cell c (row-major, row follows x)
void ref_cell_counts(int* counts) {
    const long ncell = (long)BinCnt * BinCnt;
    for (long c = 0; c < ncell; ++c) counts[c] = (int)Bins[c].size();
}

// CSR membership in the reference's own iteration order; indices are relative to `base`
void ref_cell_lists(const particle_t* base, int* start, int* member) {
    const long ncell = (long)BinCnt * BinCnt;
    int k = 0;
    for (long c = 0; c < ncell; ++c) {
        start[c] = k;
        for (particle_t* p : Bins[c]) member[k++] = (int)(p - base);
    }
    start[ncell] = k;
}
}
