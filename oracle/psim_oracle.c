/*
 * psim_oracle.c -- CPU restatement of the reference's per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under parallel-particle-simulation_b200/ may
 * include, link or call this file; it exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg can check the CUDA path against an independent
 * scalar implementation.  It is a restatement in plain C of the algorithm in the
 * reference's part1/serial.cpp (each function cites the lines it follows); it does
 * not share code with the product and it is not a fallback.
 *
 * Parity pin: the reference ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md section 4), so this restatement is pinned against the reference ITSELF:
 * oracle/Makefile compiles the unmodified /root/reference/part1/{serial,reference}.cpp
 * into oracle/_ref/ and tests/test_oracle_cpu.py checks this file against
 * them in memory (bit-exact cell ids / counts, bit-exact states for particles with
 * at most two in-range neighbours, <= 1e-12 relative otherwise), and against the
 * committed fixtures in tests/golden/ that gen_golden.py produced from those binaries.
 *
 * Summation order.  The reference visits the nine neighbour cells in the order
 * self, row-1, row+1, col-1, col+1, (row-1,col-1), (row-1,col+1), (row+1,col-1),
 * (row+1,col+1) (serial.cpp:107-115) and, inside a cell, in the iteration order of a
 * std::unordered_set<particle_t*> (serial.cpp:16,96), which depends on heap addresses
 * and insertion history and is therefore not reproducible.  This oracle keeps the
 * cell order and replaces the in-cell order by ascending (x, y, index).  A floating
 * point sum of at most two terms does not depend on order, so the two agree bit for
 * bit unless a particle has three or more in-range neighbours of which two share a
 * cell.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared (no -march, no FMA: the
 * x86-64 reference build rounds every product and sum separately).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/psim_common.h"

#define ORC_CUTOFF2 (PSIM_CUTOFF * PSIM_CUTOFF)
#define ORC_MINR2 (PSIM_MIN_R * PSIM_MIN_R)

/* serial.cpp:78 -- cells per side */
int orc_bin_count(double size) { return (int)ceil(size / PSIM_BIN_SIZE); }

/* serial.cpp:41-43, 84-86 -- row follows x, column follows y, row-major index.
 * The division is an IEEE double division: x*100.0 is not bit-identical. */
static inline int orc_axis_cell(double v) { return (int)floor(v / PSIM_BIN_SIZE); }

void orc_cell_ids(const particle_t *p, int n, int bincnt, int *cell) {
    for (int i = 0; i < n; ++i)
        cell[i] = orc_axis_cell(p[i].x) * bincnt + orc_axis_cell(p[i].y);
}

/* serial.cpp:82-87 -- per-cell population (the sizes of the reference's sets) */
void orc_cell_counts(const particle_t *p, int n, int bincnt, int *count) {
    memset(count, 0, sizeof(int) * (size_t)bincnt * (size_t)bincnt);
    for (int i = 0; i < n; ++i)
        count[orc_axis_cell(p[i].x) * bincnt + orc_axis_cell(p[i].y)] += 1;
}

/* membership lists in CSR form: start[c]..start[c+1] indexes `member`, members of
 * a cell ordered by (x, y, index).  Replaces the reference's array of hash sets. */
static const particle_t *g_sort_base;
static int orc_member_cmp(const void *a, const void *b) {
    int ia = *(const int *)a, ib = *(const int *)b;
    const particle_t *pa = g_sort_base + ia, *pb = g_sort_base + ib;
    if (pa->x != pb->x) return pa->x < pb->x ? -1 : 1;
    if (pa->y != pb->y) return pa->y < pb->y ? -1 : 1;
    return ia < ib ? -1 : (ia > ib);
}

void orc_cell_lists(const particle_t *p, int n, int bincnt, int *start, int *member) {
    size_t ncell = (size_t)bincnt * (size_t)bincnt;
    int *cell = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    memset(start, 0, sizeof(int) * (ncell + 1));
    orc_cell_ids(p, n, bincnt, cell);
    for (int i = 0; i < n; ++i) start[cell[i] + 1] += 1;
    for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    int *fill = (int *)calloc(ncell ? ncell : 1, sizeof(int));
    for (int i = 0; i < n; ++i) member[start[cell[i]] + fill[cell[i]]++] = i;
    g_sort_base = p;
    for (size_t c = 0; c < ncell; ++c) {
        int k = start[c + 1] - start[c];
        if (k > 1) qsort(member + start[c], (size_t)k, sizeof(int), orc_member_cmp);
    }
    free(fill);
    free(cell);
}

/* serial.cpp:19-36 -- one-sided short-range repulsion of `nb` on `self` */
static inline void orc_apply_force(particle_t *self, const particle_t *nb) {
    double dx = nb->x - self->x;
    double dy = nb->y - self->y;
    double r2 = dx * dx + dy * dy;
    if (r2 > ORC_CUTOFF2) return;
    r2 = fmax(r2, ORC_MINR2);
    double r = sqrt(r2);
    double coef = (1 - PSIM_CUTOFF / r) / r2 / PSIM_MASS;
    self->ax += coef * dx;
    self->ay += coef * dy;
}

/* serial.cpp:90-99 -- bounds check, then every member of the neighbour cell */
static inline void orc_interact(particle_t *p, particle_t *self, int row, int col, int bincnt,
                                const int *start, const int *member) {
    if (row < 0 || row >= bincnt || col < 0 || col >= bincnt) return;
    int c = row * bincnt + col;
    for (int k = start[c]; k < start[c + 1]; ++k) orc_apply_force(self, p + member[k]);
}

/* serial.cpp:102-117 and 121-125 -- accelerations of every particle */
void orc_compute_forces(particle_t *p, int n, double size) {
    int bincnt = orc_bin_count(size);
    size_t ncell = (size_t)bincnt * (size_t)bincnt;
    int *start = (int *)malloc(sizeof(int) * (ncell + 1));
    int *member = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    orc_cell_lists(p, n, bincnt, start, member);
    for (int row = 0; row < bincnt; ++row)
        for (int col = 0; col < bincnt; ++col) {
            int c = row * bincnt + col;
            for (int k = start[c]; k < start[c + 1]; ++k) {
                particle_t *s = p + member[k];
                s->ax = s->ay = 0;
                orc_interact(p, s, row, col, bincnt, start, member);
                orc_interact(p, s, row - 1, col, bincnt, start, member);
                orc_interact(p, s, row + 1, col, bincnt, start, member);
                orc_interact(p, s, row, col - 1, bincnt, start, member);
                orc_interact(p, s, row, col + 1, bincnt, start, member);
                orc_interact(p, s, row - 1, col - 1, bincnt, start, member);
                orc_interact(p, s, row - 1, col + 1, bincnt, start, member);
                orc_interact(p, s, row + 1, col - 1, bincnt, start, member);
                orc_interact(p, s, row + 1, col + 1, bincnt, start, member);
            }
        }
    free(member);
    free(start);
}

/* serial.cpp:46-61 -- integrate and bounce (the set bookkeeping of :64-70 is implied
 * by rebuilding the lists next step) */
void orc_move(particle_t *p, int n, double size) {
    for (int i = 0; i < n; ++i) {
        particle_t *q = p + i;
        q->vx += q->ax * PSIM_DT;
        q->vy += q->ay * PSIM_DT;
        q->x += q->vx * PSIM_DT;
        q->y += q->vy * PSIM_DT;
        while (q->x < 0 || q->x > size) {
            q->x = q->x < 0 ? -q->x : 2 * size - q->x;
            q->vx = -q->vx;
        }
        while (q->y < 0 || q->y > size) {
            q->y = q->y < 0 ? -q->y : 2 * size - q->y;
            q->vy = -q->vy;
        }
    }
}

/* serial.cpp:76-88 -- nothing persistent is needed: lists are rebuilt per step */
void orc_init_simulation(particle_t *p, int n, double size) {
    (void)p; (void)n; (void)size;
}

/* serial.cpp:119-131 */
void orc_simulate_one_step(particle_t *p, int n, double size) {
    orc_compute_forces(p, n, size);
    orc_move(p, n, size);
}

void orc_simulate_steps(particle_t *p, int n, double size, int steps) {
    for (int s = 0; s < steps; ++s) orc_simulate_one_step(p, n, size);
}

/*
 * Validation statistics (the reference has no such code, the
 * definitions are this harness's): over ordered pairs (i, j != i) with r <= cutoff,
 * dmin = min r/cutoff, davg = mean r/cutoff; plus kinetic energy and max speed.
 * out[0]=dmin out[1]=davg out[2]=ordered in-range pairs out[3]=particles with >=1
 * in-range neighbour out[4]=kinetic energy out[5]=max |v| out[6]=max in-range
 * neighbours of one particle.
 */
void orc_stats(const particle_t *p, int n, double size, double *out) {
    int bincnt = orc_bin_count(size);
    size_t ncell = (size_t)bincnt * (size_t)bincnt;
    int *start = (int *)malloc(sizeof(int) * (ncell + 1));
    int *member = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    orc_cell_lists(p, n, bincnt, start, member);
    double dmin = 1.0, dsum = 0.0, ke = 0.0, vmax = 0.0;
    long long pairs = 0, touched = 0;
    int maxnb = 0;
    for (int i = 0; i < n; ++i) {
        int row = orc_axis_cell(p[i].x), col = orc_axis_cell(p[i].y), nb = 0;
        for (int dr = -1; dr <= 1; ++dr)
            for (int dc = -1; dc <= 1; ++dc) {
                int r = row + dr, c = col + dc;
                if (r < 0 || r >= bincnt || c < 0 || c >= bincnt) continue;
                int cell = r * bincnt + c;
                for (int k = start[cell]; k < start[cell + 1]; ++k) {
                    int j = member[k];
                    if (j == i) continue;
                    double dx = p[j].x - p[i].x, dy = p[j].y - p[i].y;
                    double r2 = dx * dx + dy * dy;
                    if (r2 > ORC_CUTOFF2) continue;
                    double d = sqrt(r2) / PSIM_CUTOFF;
                    if (d < dmin) dmin = d;
                    dsum += d;
                    ++pairs;
                    ++nb;
                }
            }
        if (nb) ++touched;
        if (nb > maxnb) maxnb = nb;
        double v2 = p[i].vx * p[i].vx + p[i].vy * p[i].vy;
        ke += 0.5 * PSIM_MASS * v2;
        if (sqrt(v2) > vmax) vmax = sqrt(v2);
    }
    out[0] = dmin;
    out[1] = pairs ? dsum / (double)pairs : 0.0;
    out[2] = (double)pairs;
    out[3] = (double)touched;
    out[4] = ke;
    out[5] = vmax;
    out[6] = (double)maxnb;
    free(member);
    free(start);
}
