/*
 * psim_common.h -- the simulation-kernel interface this library drops in behind.
 *
 * This is the contract of the reference's `common.h` (reference part1/common.h:5-25,
 * byte-identical in part3/common.h): the 48-byte AoS particle record, the seven
 * physics constants and the two C++-linkage entry points that the reference drivers
 * (part1/main.cpp:122,129 and part3/main.cu:127,130) call.  A translation unit may
 * include either the reference's own common.h or this header; both describe the
 * same ABI (struct layout and the mangled names
 * _Z15init_simulationP10particle_tid / _Z17simulate_one_stepP10particle_tid).
 *
 * The PSIM_* spellings are what this repository's own sources use; the lower-case
 * macro names of the reference are only provided on request (PSIM_REFERENCE_MACROS)
 * because names such as `dt` and `mass` collide with ordinary identifiers.
 */
#ifndef PSIM_COMMON_H
#define PSIM_COMMON_H

/* physics / run constants (reference part1/common.h:5-11) */
#define PSIM_NSTEPS    1000
#define PSIM_SAVEFREQ  10
#define PSIM_DENSITY   0.0005
#define PSIM_MASS      0.01
#define PSIM_CUTOFF    0.01
#define PSIM_MIN_R     (PSIM_CUTOFF / 100)
#define PSIM_DT        0.0005
/* side of one cutoff cell (reference part1/serial.cpp:11, part3/gpu.cu:12) */
#define PSIM_BIN_SIZE  0.01

#ifdef PSIM_REFERENCE_MACROS
#define nsteps   PSIM_NSTEPS
#define savefreq PSIM_SAVEFREQ
#define density  PSIM_DENSITY
#define mass     PSIM_MASS
#define cutoff   PSIM_CUTOFF
#define min_r    PSIM_MIN_R
#define dt       PSIM_DT
#endif

/* particle record: six doubles, 48 bytes, array-of-structures, caller owned
 * (reference part1/common.h:14-21) */
typedef struct particle_t {
    double x, y;    /* position                    */
    double vx, vy;  /* velocity                    */
    double ax, ay;  /* acceleration of the last step */
} particle_t;

#ifdef __cplusplus
/* The reference declares these with C++ linkage (no extern "C"), so the drop-in
 * shim (csrc/psim_shim.cpp) must export the C++-mangled names.
 * `parts` may be a host pointer (part1/main.cpp) or a device pointer (part3/main.cu). */
void init_simulation(particle_t* parts, int num_parts, double size);
void simulate_one_step(particle_t* parts, int num_parts, double size);
#endif

#endif /* PSIM_COMMON_H */
