/*
 * psim.h -- C ABI of libpsim: the B200 (sm_100a) implementation of the reference's per-timestep
 * hot path (cell binning -> 3x3-neighbourhood short-range force -> move with reflective walls).
 *
 * Every entry point is extern "C", takes plain pointers and sizes (no C++ or torch types) and
 * returns an int status (PSIM_OK == 0).  The reference has no error channel (void functions; its
 * CUDA part prints "GPUassert: ..." and exits, reference part3/gpu.cu:70-77); the C++ shim
 * (csrc/psim_shim.cpp) that exports the reference's two mangled C++ names maps a non-zero status
 * to exactly that behaviour.
 *
 * There is NO CPU fallback: every function that computes needs a CUDA device and fails with
 * PSIM_ERR_NO_DEVICE / PSIM_ERR_CUDA otherwise.
 *
 * Reference interface replaced by each group (paths relative to the reference repository):
 *   psim_create            init_simulation      part1/serial.cpp:76-88, part1/openmp.cpp:42-66,
 *                                               part3/gpu.cu:174-185   (+ the H2D upload of
 *                                               part3/main.cu:120-122 when `parts` is a host pointer)
 *   psim_step              simulate_one_step    part1/serial.cpp:119-131, part1/openmp.cpp:69-83,
 *                                               part3/gpu.cu:187-208   (n calls fused into one batch)
 *   psim_read_particles    the driver's view of `parts` after a step: part1/main.cpp:135-136 (host),
 *                                               part3/main.cu:134-136 (device array + cudaMemcpy D2H)
 *   psim_read_cells        the bin structure    part1/serial.cpp:14-16,41-43,84-86;
 *                                               part3/gpu.cu:92-112 (Bins / Bin_Sizes)
 *   psim_gather            gather_for_save      part2/common.h:32, part2/mpi.cpp:371-402 (MPI flavour; the 5-argument
 *                                               entry points themselves are exported by csrc/psim_mpi_shim.cpp)
 *   psim_stats             (no reference code; validation statistics of SURVEY.md section 8c-5)
 *   psim_init_particles    init_particles       part1/main.cpp:31-59 (bit-compatible host replay)
 *   psim_generate_particles_device               the same construction in parallel on the device (alternative seeding mode)
 *   psim_save_frame        save                 part1/main.cpp:15-28 (same bytes; std::to_chars formatter,
 *                                               one fwrite per 64 k particles instead of a flush per line)
 */
#ifndef PSIM_H
#define PSIM_H

#include <stddef.h>

#include "psim_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
#define PSIM_OK               0
#define PSIM_ERR_INVALID      1  /* bad argument / bad handle                                   */
#define PSIM_ERR_NO_DEVICE    2  /* no usable CUDA device (there is no CPU path)                */
#define PSIM_ERR_CUDA         3  /* a CUDA runtime call or kernel failed                        */
#define PSIM_ERR_CAPACITY     4  /* a tile / halo list / migration outbox overflowed on device  */
#define PSIM_ERR_STATE        5  /* call not valid in the current state                         */
#define PSIM_ERR_COMM         6  /* NCCL / multi-GPU exchange failure                           */
#define PSIM_ERR_UNSUPPORTED  7  /* configuration the selected engine cannot run                */

/* ---- engines ---- */
#define PSIM_ENGINE_AUTO      0  /* kstep when the particle density fits its tiles, else cellsort */
#define PSIM_ENGINE_CELLSORT  1  /* per step: atomic histogram over cutoff cells, single-pass
                                    exclusive scan, scatter into cell-sorted SoA, force+move     */
#define PSIM_ENGINE_TILED     2  /* persistent tile-resident SoA; one fused kernel per step does
                                    in-shared-memory binning, force, move, re-tiling, halo export */
#define PSIM_ENGINE_KSTEP     3  /* persistent tile-resident SoA; one kernel advances a tile and its
                                    5-cell halo K (<= 4) steps in shared memory -- binning, force,
                                    move every step -- and re-tiles once per launch; results are
                                    bit-identical to stepping one step at a time                    */

/* flags for psim_step */
#define PSIM_STEP_DEFAULT     0  /* accelerations are materialised for the LAST step of the batch */
#define PSIM_STEP_ACCEL_ALL   1  /* store ax, ay after every step of the batch                    */
#define PSIM_STEP_ACCEL_NONE  2  /* never store ax, ay (a read-back then returns ax = ay = 0)      */

typedef struct psim_sim psim_sim; /* opaque */

typedef struct psim_config {
    int   engine;        /* PSIM_ENGINE_*                                                        */
    int   device;        /* CUDA device ordinal; -1 = the calling thread's current device        */
    void* stream;        /* cudaStream_t to run on; NULL = a private non-blocking stream         */
    int   tile_cells;    /* tiled / kstep engines: cutoff cells per tile side (16, 32 or 64; kstep also 48); 0 = auto */
    int   steps_per_launch; /* kstep engine: time steps fused per kernel launch, 1 .. 3; 0 = default (3, or PSIM_KSTEPS) */
    /* 1-D slab decomposition.  nranks == 1: the whole box.               */
    int   rank;          /* this slab's index along x (cell rows)                                */
    int   nranks;        /* number of slabs                                                      */
    int   reserved[8];
} psim_config;

typedef struct psim_stats_t {
    double dmin;           /* min r/cutoff over ordered in-range pairs (1.0 if none)            */
    double davg;           /* mean r/cutoff over ordered in-range pairs                         */
    double kinetic_energy; /* sum 0.5 m v^2                                                     */
    double vmax;           /* max |v|                                                           */
    long long pairs;       /* ordered in-range pairs (i, j != i)                                */
    long long touched;     /* particles with >= 1 in-range neighbour                            */
    int    max_neighbours; /* max in-range neighbours of one particle                           */
    int    max_cell_count; /* max population of one cutoff cell                                 */
} psim_stats_t;

typedef struct psim_info_t {
    int engine;            /* engine actually selected                                          */
    int bin_count;         /* cutoff cells per box side = ceil(size / 0.01)                     */
    int tile_cells;        /* tiled engine: cells per tile side                                 */
    int tiles_per_side;
    int tile_capacity;     /* particle slots per tile                                           */
    int device;
    int num_parts;         /* particles in the box; for a slab (nranks > 1): particles it owned at the
                              last observation call (read / stats / hash / gather)                */
    int rank, nranks;
    int row_begin, row_end;/* cell rows [begin, end) owned by this slab                         */
    long long steps_done;
    long long kernel_launches; /* kernels this handle launched so far                           */
    long long device_bytes;    /* device memory held                                            */
    /* tiled engine, as of the last psim_sync / observation call: high-water marks against the
     * fixed capacities (outbox records per tile-step, edge halo list length, tile population, apron) */
    int hw_leavers, hw_halo_list, hw_tile_population, hw_apron;
    int outbox_capacity, halo_list_capacity;
    int reserved_hw_pairs;  /* tiled engine: most candidate pairs one tile listed in a step (capacity is internal) */
    /* kstep engine */
    int halo_cells;        /* halo width H in cutoff cells                                       */
    int steps_per_launch;  /* K: time steps fused per kernel launch (PSIM_KSTEPS, default 4)     */
    int region_capacity;   /* particles of one tile + halo region the kernel can hold            */
    int recoveries;        /* batches replayed one step per launch after a speed-bound violation  */
    int input_on_device;   /* psim_create was handed a device pointer (part3/main.cu flavour)    */
    int engine_switches;   /* kstep -> cellsort hand-overs (stripe / region overflow, or a particle faster than
                              the one-step halo bound): the handle keeps running on the cellsort engine */
} psim_info_t;

/* ---- errors ---- */
const char* psim_error_string(int status);
/* detail of the last failure on the calling thread ("" if none) */
const char* psim_last_error(void);

/* ---- lifecycle ---- */
/* Create the CUDA context on `device` (-1: current) ahead of time.  The reference's CUDA driver pays this cost before
 * its timer starts (its cudaMalloc at part3/main.cu:120-122 precedes the clock at :125); a host-pointer driver would
 * otherwise pay it inside init_simulation.  Optional: psim_create does it implicitly. */
int psim_device_init(int device);
/* number of visible CUDA devices (the multi-process shim spreads its ranks over them) */
int psim_device_count(int* count);
/* Page-lock (or release) a caller-owned host array so that the read-backs into it run at full PCIe speed.  Optional;
 * the drop-in shim does it for the driver's `parts` array, which it is handed on every call anyway. */
int psim_host_register(void* host_ptr, size_t bytes);
int psim_host_unregister(void* host_ptr);
void psim_config_default(psim_config* cfg);
/* cells per side, ceil(size / 0.01): reference part1/serial.cpp:78 */
int  psim_bin_count(double size);

/* Build the simulation state from `num_parts` AoS records.  `parts` may be a host pointer
 * (part1/main.cpp flavour) or a device pointer (part3/main.cu flavour); the kind is detected.
 * ax, ay of the input are ignored (the reference driver leaves them uninitialised).
 * With nranks > 1 the slab keeps only the particles whose cell row falls in its row range.  For a HOST array and
 * nranks > 1 the upload is cooperative and completes in psim_comm_connect: every rank uploads 1/nranks of the array
 * (all ranks must be handed the same array, as the reference's driver broadcasts it, part2/main.cpp:150) and the records
 * travel to their slabs GPU to GPU -- `parts` must stay valid and unchanged until psim_comm_connect has returned, and
 * capacity errors of the initial tiling are reported there (PSIM_COOP_UPLOAD=0: every rank uploads everything in
 * psim_create). */
int psim_create(psim_sim** out, const psim_config* cfg, const particle_t* parts, int num_parts, double size);
int psim_destroy(psim_sim* sim);

/* ---- stepping ---- */
/* Enqueue `nsteps` time steps on the handle's stream (asynchronous). */
int psim_step(psim_sim* sim, int nsteps, int flags);
/* Wait for the stream and report sticky device-side errors (capacity overflow etc.). */
int psim_sync(psim_sim* sim);

/* ---- observation (all of these synchronise) ---- */
/* Write the current state, in ORIGINAL particle order, to `dst` (host or device pointer, the
 * caller's full array of `num_parts_total` records; a slab only writes the particles it owns).
 * Fields: x y vx vy of the current step; ax ay = the acceleration used by the LAST step if that step
 * stored it (see PSIM_STEP_*), else 0 -- every engine. */
int psim_read_particles(psim_sim* sim, particle_t* dst);
/* positions only: xy[2*i] = x_i, xy[2*i+1] = y_i, host or device pointer */
int psim_read_positions(psim_sim* sim, double* xy);
/* The same, asynchronously (the save path, reference part1/main.cpp:135-136 / part3/main.cu:134-136): _begin enqueues the
 * original-order gather behind the steps enqueued so far and the device->host copy on a separate stream, then returns;
 * steps enqueued afterwards run while the copy flies.  _end blocks until the oldest un-awaited read has landed in its
 * host buffer (which should be page-locked: psim_host_register).  At most two reads in flight.  The state that is read
 * is checked for device-side errors at the next synchronising call. */
int psim_read_positions_begin(psim_sim* sim, double* xy_host);
int psim_read_positions_end(psim_sim* sim);
/* cell_of_particle[i] = row*bin_count+col with row = floor(x/0.01), col = floor(y/0.01)
 * (bit-exact IEEE division, reference part1/serial.cpp:41-43); cell_counts[c] = population of
 * cell c (bin_count^2 ints).  Either pointer may be NULL.  HOST pointers. */
int psim_read_cells(psim_sim* sim, int* cell_of_particle, int* cell_counts);
/* membership lists in CSR form (the reference's Bins): cell_start has bin_count^2+1 entries,
 * members has num_parts entries, members of a cell in ascending original index.  HOST pointers. */
int psim_read_cell_lists(psim_sim* sim, int* cell_start, int* members);
int psim_stats(psim_sim* sim, psim_stats_t* out);
/* Fingerprint of the particles this handle owns: the sum (mod 2^64) over particles of a 64-bit mix of the original
 * index and the bits of x, y, vx, vy -- independent of storage order, tile size, engine and slab count, so the sums of
 * all slabs of one run add up to the same number as the single-GPU run (bench.py check.state_hash).  `owned` (may be
 * NULL) receives the number of particles hashed. */
int psim_state_hash(psim_sim* sim, unsigned long long* hash, long long* owned);
int psim_info(psim_sim* sim, psim_info_t* out);

/* ---- driver helpers (host side) ---- */
/* The reference driver's particle generator: std::mt19937(seed), shuffled lattice, float
 * velocities in [-1,1) -- reference part1/main.cpp:31-59.  seed 0 = std::random_device.
 * ax, ay are set to 0. */
int psim_init_particles(particle_t* parts_host, int num_parts, double size, int seed);
/* Alternative seeding mode for large N: the same construction (one particle per site of the reference's sx x sy lattice,
 * sites handed out by a pseudo-random permutation, float velocities uniform in [-1, 1), ax = ay = 0) generated IN PARALLEL
 * on the device into a device array -- no host generator, no host->device upload (at 160 M particles the sequential
 * reference generator and a 7.7 GB upload dominate everything else).  Deterministic in (num_parts, seed), but NOT the
 * reference's bits for that seed: psim_init_particles is the bit-compatible replay.  Asynchronous on `stream`
 * (a cudaStream_t; NULL = the default stream). */
int psim_generate_particles_device(particle_t* parts_device, int num_parts, double size, int seed, void* stream);
/* Append one frame to an open trajectory file in the reference's text format
 * (part1/main.cpp:15-28): first call writes "N size", every call N lines "x y" + blank line.
 * `file` is a FILE*; `xy` the interleaved positions from psim_read_positions. */
int psim_save_frame(void* file, const double* xy, int num_parts, double size, int first);

/* ---- multi-GPU slab exchange (one process per GPU) ---- */
/* The slab decomposition the engine uses (precedent: reference part2/mpi.cpp:258-270, rows of cells along x split
 * contiguously over the ranks): cell rows [*row_begin, *row_end) of a box with `bin_count` cells per side belong to
 * `rank` of `nranks` when tiles are `tile_cells` cells wide (0 = the default engine's choice for that box).  Host only. */
int psim_slab_rows(int bin_count, int tile_cells, int rank, int nranks, int* row_begin, int* row_end);
/* 128-byte NCCL unique id: rank 0 creates it, the launcher broadcasts it, every rank connects. */
int psim_comm_unique_id(unsigned char id128[128]);
int psim_comm_connect(psim_sim* sim, const unsigned char id128[128]);
/* Collective over the slabs (every rank calls it): rank `root` receives ALL particles, in original order, in `dst`
 * (host or device array of num_parts_total records; ignored on the other ranks).  The reference's gather_for_save
 * (part2/common.h:32, part2/mpi.cpp:371-402: every rank sends its particles to rank 0, which places them by id); here the
 * records travel GPU to GPU over NCCL and are placed by a kernel. */
int psim_gather(psim_sim* sim, particle_t* dst, int root);

#ifdef __cplusplus
}
#endif
#endif /* PSIM_H */
