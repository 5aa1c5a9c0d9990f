// psim_capi.cu -- the extern "C" boundary (include/psim.h): handle life cycle, engine dispatch,
// observation kernels (original-order read-back, cell ids / counts / lists, statistics).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <thread>

#include "psim_device.cuh"
#include "psim_internal.h"

namespace psim {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int fail(int status, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return status;
}

// ------------------------------------------------------------------------------------------
// observation kernels
// ------------------------------------------------------------------------------------------
constexpr int kObsThreads = 256;

// compact SoA -> caller's AoS in original order (reference part1/common.h:14-21 record)
__global__ void __launch_bounds__(kObsThreads) soa_to_aos_kernel(SoAView v, bool have_acc, particle_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    double2* q = reinterpret_cast<double2*>(out + v.id[i]);
    q[0] = make_double2(v.x[i], v.y[i]);
    q[1] = make_double2(v.vx[i], v.vy[i]);
    q[2] = have_acc ? make_double2(v.ax[i], v.ay[i]) : make_double2(0.0, 0.0);
}

__global__ void __launch_bounds__(kObsThreads) soa_to_xy_kernel(SoAView v, double2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    out[v.id[i]] = make_double2(v.x[i], v.y[i]);
}

// compact records for the multi-slab host path: rec[i] = {x y vx vy ax ay}
__global__ void __launch_bounds__(kObsThreads) soa_pack_kernel(SoAView v, bool have_acc, particle_t* __restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    double2* q = reinterpret_cast<double2*>(rec + i);
    q[0] = make_double2(v.x[i], v.y[i]);
    q[1] = make_double2(v.vx[i], v.vy[i]);
    q[2] = have_acc ? make_double2(v.ax[i], v.ay[i]) : make_double2(0.0, 0.0);
}

// gathered records (any order) + their original indices -> AoS in original order
__global__ void __launch_bounds__(kObsThreads) place_records_kernel(const particle_t* __restrict__ rec, const int* __restrict__ ids, int n,
                                                                    particle_t* __restrict__ out) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;   // three lanes per 48-byte record
    const int i = u / 3, piece = u - 3 * i;
    if (i >= n) return;
    reinterpret_cast<double2*>(out + ids[i])[piece] = reinterpret_cast<const double2*>(rec + i)[piece];
}

__global__ void __launch_bounds__(kObsThreads) cell_of_particle_kernel(SoAView v, int bincnt, int* __restrict__ cell_of) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    cell_of[v.id[i]] = axis_cell(v.x[i], bincnt) * bincnt + axis_cell(v.y[i], bincnt);
}

// members in arrival order, then ranked by original index inside each cell
__global__ void __launch_bounds__(kObsThreads) members_raw_kernel(SoAView v, int bincnt, const int* __restrict__ cell_start,
                                                                  const int* __restrict__ slot, int* __restrict__ raw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const int c = axis_cell(v.x[i], bincnt) * bincnt + axis_cell(v.y[i], bincnt);
    raw[cell_start[c] + slot[i]] = v.id[i];
}
__global__ void __launch_bounds__(kObsThreads) members_rank_kernel(SoAView v, int bincnt, const int* __restrict__ cell_start,
                                                                   const int* __restrict__ raw, int* __restrict__ members) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const int c = axis_cell(v.x[i], bincnt) * bincnt + axis_cell(v.y[i], bincnt);
    const int s = cell_start[c], e = cell_start[c + 1], me = v.id[i];
    int rank = 0;
    for (int k = s; k < e; ++k) rank += raw[k] < me;
    members[s + rank] = me;
}

// sorted copies of the positions for the statistics pass
__global__ void __launch_bounds__(kObsThreads) sort_xy_kernel(SoAView v, int bincnt, const int* __restrict__ cell_start,
                                                              const int* __restrict__ slot, double* __restrict__ sx,
                                                              double* __restrict__ sy, unsigned char* __restrict__ owned_sorted,
                                                              int n_owned) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const int c = axis_cell(v.x[i], bincnt) * bincnt + axis_cell(v.y[i], bincnt);
    const int d = cell_start[c] + slot[i];
    sx[d] = v.x[i];
    sy[d] = v.y[i];
    if (owned_sorted) owned_sorted[d] = i < n_owned;   // entries past n_owned are ghost particles: neighbours only
}

// order-independent fingerprint of the owned particles: sum over particles of a 64-bit mix of (id, x, y, vx, vy) bits
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(kObsThreads) state_hash_kernel(SoAView v, unsigned long long* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = 0;
    if (i < v.n) {
        h = mix64(((unsigned long long)(unsigned)v.id[i] + 1ull) * 0x9E3779B97F4A7C15ull);
        h = mix64(h ^ (unsigned long long)__double_as_longlong(v.x[i]));
        h = mix64(h ^ (unsigned long long)__double_as_longlong(v.y[i]));
        h = mix64(h ^ (unsigned long long)__double_as_longlong(v.vx[i]));
        h = mix64(h ^ (unsigned long long)__double_as_longlong(v.vy[i]));
    }
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}

// ---- device-side particle generation ------------------------------------------------------------------------
// The reference generator (part1/main.cpp:31-59) is a sequential Fisher-Yates draw from one mt19937 stream with
// rejection-sampled integers: it cannot be replayed in parallel bit for bit (psim_init_particles is its host replay).  This is
// the documented ALTERNATIVE seeding mode with the same construction -- every particle on its own site of the same
// sx x sy lattice, sites assigned by a pseudo-random permutation, float velocities uniform in [-1, 1) -- computed per particle:
// the permutation is a 4-round Feistel network on the smallest even-bit domain >= N with cycle walking, keyed by the seed;
// velocities come from a counter-based hash of (seed, i).  Same statistics, different bits than the reference for a given seed.
__device__ __forceinline__ unsigned feistel_round(unsigned r, unsigned key, int half_bits) {
    unsigned long long z = ((unsigned long long)r + 1ull) * 0x9E3779B97F4A7C15ull + key;
    z = (z ^ (z >> 29)) * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 32;
    return (unsigned)z & ((1u << half_bits) - 1u);
}
__global__ void __launch_bounds__(kObsThreads) generate_particles_kernel(particle_t* __restrict__ out, int n, double size, unsigned seed,
                                                                        int sx, int sy, int half_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned mask = (1u << half_bits) - 1u;
    unsigned v = (unsigned)i;
    do {   // cycle walking: the Feistel network permutes [0, 4^half_bits); follow the cycle until it re-enters [0, n)
        unsigned l = v >> half_bits, r = v & mask;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned t = l ^ feistel_round(r, seed * 0x85EBCA6Bu + (unsigned)k * 0xC2B2AE35u, half_bits);
            l = r;
            r = t;
        }
        v = (l << half_bits) | r;
    } while (v >= (unsigned)n);
    const int site = (int)v;
    unsigned long long h = ((unsigned long long)seed << 32 | (unsigned)i) * 0xD6E8FEB86659FD93ull;
    h = (h ^ (h >> 32)) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    const float vx = (float)((unsigned)(h >> 40)) * (1.0f / 8388608.0f) - 1.0f;          // 24 bits -> [-1, 1)
    const float vy = (float)((unsigned)(h >> 8) & 0xFFFFFFu) * (1.0f / 8388608.0f) - 1.0f;
    double2* q = reinterpret_cast<double2*>(out + i);
    q[0] = make_double2(size * (1. + (site % sx)) / (1 + sx), size * (1. + (site / sx)) / (1 + sy));   // reference main.cpp:49-50
    q[1] = make_double2((double)vx, (double)vy);
    q[2] = make_double2(0.0, 0.0);
}

struct StatsPartial {
    double dmin, dsum, ke, vmax;
    long long pairs, touched;
    int maxnb, maxcell;
};

__global__ void __launch_bounds__(kObsThreads) stats_kernel(const double* __restrict__ sx, const double* __restrict__ sy, int n,
                                                            SoAView v, int bincnt, const int* __restrict__ cell_start,
                                                            const unsigned char* __restrict__ owned_sorted, int n_owned,
                                                            StatsPartial* __restrict__ partial) {
    __shared__ StatsPartial s_part[kObsThreads / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    StatsPartial p{1.0, 0.0, 0.0, 0.0, 0, 0, 0, 0};
    if (i < n && (!owned_sorted || owned_sorted[i])) {
        const double xi = sx[i], yi = sy[i];
        const int row = axis_cell(xi, bincnt), col = axis_cell(yi, bincnt);
        const int c_lo = max(col - 1, 0), c_hi = min(col + 1, bincnt - 1);
        int nb = 0;
        for (int dr = -1; dr <= 1; ++dr) {
            const int rr = row + dr;
            if (rr < 0 || rr >= bincnt) continue;
            const long long b = (long long)rr * bincnt;
            for (int k = cell_start[b + c_lo]; k < cell_start[b + c_hi + 1]; ++k) {
                if (k == i) continue;
                const double dx = __dsub_rn(sx[k], xi), dy = __dsub_rn(sy[k], yi);
                const double r2 = pair_r2(dx, dy);
                if (r2 > kCutoff2) continue;
                const double d = __ddiv_rn(__dsqrt_rn(r2), kCutoff);
                p.dmin = fmin(p.dmin, d);
                p.dsum += d;
                ++nb;
            }
        }
        p.pairs = nb;
        p.touched = nb > 0;
        p.maxnb = nb;
        const long long c = (long long)row * bincnt + col;
        p.maxcell = cell_start[c + 1] - cell_start[c];
    }
    if (i < n_owned) {
        const double vx = v.vx[i], vy = v.vy[i];  // any order: kinetic terms are per particle
        const double v2 = vx * vx + vy * vy;
        p.ke = 0.5 * kMass * v2;
        p.vmax = sqrt(v2);
    }
    // warp, then block reduction
    for (int o = 16; o > 0; o >>= 1) {
        p.dmin = fmin(p.dmin, __shfl_xor_sync(0xffffffffu, p.dmin, o));
        p.dsum += __shfl_xor_sync(0xffffffffu, p.dsum, o);
        p.ke += __shfl_xor_sync(0xffffffffu, p.ke, o);
        p.vmax = fmax(p.vmax, __shfl_xor_sync(0xffffffffu, p.vmax, o));
        p.pairs += __shfl_xor_sync(0xffffffffu, p.pairs, o);
        p.touched += __shfl_xor_sync(0xffffffffu, p.touched, o);
        p.maxnb = max(p.maxnb, __shfl_xor_sync(0xffffffffu, p.maxnb, o));
        p.maxcell = max(p.maxcell, __shfl_xor_sync(0xffffffffu, p.maxcell, o));
    }
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        StatsPartial t = s_part[0];
        for (int w = 1; w < kObsThreads / 32; ++w) {
            const StatsPartial& q = s_part[w];
            t.dmin = fmin(t.dmin, q.dmin);
            t.dsum += q.dsum;
            t.ke += q.ke;
            t.vmax = fmax(t.vmax, q.vmax);
            t.pairs += q.pairs;
            t.touched += q.touched;
            t.maxnb = max(t.maxnb, q.maxnb);
            t.maxcell = max(t.maxcell, q.maxcell);
        }
        partial[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
static bool pointer_on_device(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// A pinned (page-locked, mapped) host array can be written by a kernel directly: returns the device alias of p, or null.
static void* device_alias_of_host(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

static int view_of(psim_sim* sim, SoAView* v, bool* have_acc) {
    if (sim->engine == PSIM_ENGINE_CELLSORT) {
        PSIM_TRY(cellsort_view(sim, v));
        *have_acc = cellsort_acc_valid(sim);   // zeros instead of accelerations that belong to an older particle order
    } else if (sim->engine == PSIM_ENGINE_KSTEP) {
        PSIM_TRY(kstep_view(sim, v));
        *have_acc = true;  // zeros are substituted when the stored accelerations are stale
    } else {
        PSIM_TRY(tiled_view(sim, v));
        *have_acc = true;  // the gather already substitutes zeros when accelerations are stale
    }
    sim->owned_last = v->n;
    return PSIM_OK;
}

static int engine_writeback(psim_sim* sim, particle_t* d_out, double2* d_xy);

// kstep -> cellsort hand-over (see check_device_error): the kstep state (rewound to the last good launch) is written out in
// original order, the cellsort engine is built from it and runs the steps the failed launches still owed.
static int switch_to_cellsort(psim_sim* sim, int pending_steps, bool store_last) {
    cudaStream_t s = sim->stream;
    DeviceArena tmp;
    particle_t* aos = nullptr;
    PSIM_TRY(tmp.alloc(&aos, (size_t)std::max(sim->n_total, 1)));
    int st = kstep_writeback(sim, aos, nullptr);
    if (st == PSIM_OK && cudaStreamSynchronize(s) != cudaSuccess) st = fail(PSIM_ERR_CUDA, "engine hand-over: %s", cudaGetErrorString(cudaGetLastError()));
    if (st == PSIM_OK) {
        kstep_destroy(sim);
        if (cudaMemsetAsync(sim->d_err, 0, kErrWords * sizeof(int), s) != cudaSuccess) st = fail(PSIM_ERR_CUDA, "engine hand-over: memset");
    }
    if (st == PSIM_OK) st = cellsort_create(sim, aos, sim->n_total);
    if (st == PSIM_OK && cudaStreamSynchronize(s) != cudaSuccess) st = fail(PSIM_ERR_CUDA, "engine hand-over: %s", cudaGetErrorString(cudaGetLastError()));
    tmp.release();
    PSIM_TRY(st);
    sim->engine = PSIM_ENGINE_CELLSORT;
    ++sim->engine_switches;
    if (pending_steps > 0) PSIM_TRY(cellsort_step(sim, pending_steps, store_last ? PSIM_STEP_DEFAULT : PSIM_STEP_ACCEL_NONE));
    PSIM_CUDA(cudaMemcpyAsync(sim->h_err, sim->d_err, kErrWords * sizeof(int), cudaMemcpyDeviceToHost, s));
    PSIM_CUDA(cudaStreamSynchronize(s));
    return PSIM_OK;
}

static int check_device_error(psim_sim* sim) {
    for (int attempt = 0;; ++attempt) {
        PSIM_CUDA(cudaMemcpyAsync(sim->h_err, sim->d_err, kErrWords * sizeof(int), cudaMemcpyDeviceToHost, sim->stream));
        PSIM_CUDA(cudaStreamSynchronize(sim->stream));
        if (sim->engine != PSIM_ENGINE_KSTEP || !sim->kstep) break;
        // kstep engine: a launch that exceeded its speed bound left its input untouched and is replayed one step per launch
        bool replayed = false, pending_store = false;
        int pending_steps = -1;
        const int st = kstep_after_sync(sim, &replayed, &pending_steps, &pending_store);
        if (st != PSIM_OK && pending_steps >= 0 && sim->nranks == 1) {
            // Not recoverable inside the kstep engine (a stripe or region overflowed, or a particle is faster than even the
            // one-step halo allows), but the input of the failed launch is intact: hand the state to the cellsort engine,
            // which has no capacities and no speed limit, and run the steps that are still owed there.
            PSIM_TRY(switch_to_cellsort(sim, pending_steps, pending_store));
            break;
        }
        if (st != PSIM_OK) {
            const int e = sim->h_err[0];
            return fail(st,
                        "kstep engine: launch %d failed with device flags 0x%x:%s%s%s (high-water marks: stripe %d, region %d); the "
                        "state was rewound to the input of that launch (step %lld)",
                        sim->h_err[8] - 1, e, (e & kErrTileOverflow) ? " stripe-overflow" : "",
                        (e & kErrSmemOverflow) ? " region-overflow" : "", (e & kErrSpeedBound) ? " speed-bound" : "", sim->h_err[3],
                        sim->h_err[4], sim->steps_done);
        }
        if (!replayed) return PSIM_OK;
        if (attempt > 64) return fail(PSIM_ERR_STATE, "kstep engine: replay does not converge");
    }
    const int e = *sim->h_err;
    if (sim->h_err[5]) return fail(PSIM_ERR_STATE, "debug guard 0x%x tripped (last tile %d, flags 0x%x)", sim->h_err[5], sim->h_err[6], e);
    if (e == 0) return PSIM_OK;
    return fail(PSIM_ERR_CAPACITY,
                "device capacity error 0x%x:%s%s%s%s%s (high-water marks: leavers/tile-step %d, edge halo list %d, tile population "
                "%d, apron %d)",
                e, (e & kErrTileOverflow) ? " tile-overflow" : "", (e & kErrHaloOverflow) ? " halo-list-overflow" : "",
                (e & kErrOutboxOverflow) ? " outbox-overflow" : "", (e & kErrSmemOverflow) ? " apron-staging-overflow" : "",
                (e & kErrLostParticle) ? " particle-skipped-a-tile" : "", sim->h_err[1], sim->h_err[2], sim->h_err[3],
                sim->h_err[4]);
}

// read-back staging buffer, kept between calls (the drop-in drivers read the state back every savefreq steps)
static int ensure_scratch(psim_sim* sim, size_t bytes, void** out) {
    if (sim->scratch_bytes < bytes) {
        sim->scratch.release();
        sim->scratch_ptr = nullptr;
        sim->scratch_bytes = 0;
        char* p = nullptr;
        PSIM_TRY(sim->scratch.alloc(&p, bytes));
        sim->scratch_ptr = p;
        sim->scratch_bytes = bytes;
    }
    *out = sim->scratch_ptr;
    return PSIM_OK;
}

static int engine_writeback(psim_sim* sim, particle_t* d_out, double2* d_xy) {
    return sim->engine == PSIM_ENGINE_KSTEP ? kstep_writeback(sim, d_out, d_xy) : tiled_writeback(sim, d_out, d_xy);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace psim

using namespace psim;

extern "C" {

const char* psim_error_string(int status) {
    switch (status) {
        case PSIM_OK: return "ok";
        case PSIM_ERR_INVALID: return "invalid argument";
        case PSIM_ERR_NO_DEVICE: return "no CUDA device (libpsim has no CPU path)";
        case PSIM_ERR_CUDA: return "CUDA error";
        case PSIM_ERR_CAPACITY: return "device-side capacity overflow";
        case PSIM_ERR_STATE: return "invalid state";
        case PSIM_ERR_COMM: return "communication error";
        case PSIM_ERR_UNSUPPORTED: return "unsupported configuration";
    }
    return "unknown status";
}

const char* psim_last_error(void) { return g_last_error.c_str(); }

void psim_config_default(psim_config* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof *cfg);
    cfg->engine = PSIM_ENGINE_AUTO;
    cfg->device = -1;
    cfg->nranks = 1;
}

int psim_bin_count(double size) { return (int)std::ceil(size / PSIM_BIN_SIZE); }

int psim_host_register(void* host_ptr, size_t bytes) {
    if (!host_ptr || !bytes) return fail(PSIM_ERR_INVALID, "psim_host_register: NULL / empty");
    if (pointer_on_device(host_ptr)) return PSIM_OK;   // nothing to pin
    const cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return PSIM_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PSIM_ERR_CUDA, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    return PSIM_OK;
}

int psim_host_unregister(void* host_ptr) {
    if (!host_ptr) return PSIM_OK;
    if (cudaHostUnregister(host_ptr) != cudaSuccess) cudaGetLastError();
    return PSIM_OK;
}

int psim_generate_particles_device(particle_t* parts_device, int num_parts, double size, int seed, void* stream) {
    if (num_parts < 0 || (num_parts > 0 && !parts_device)) return fail(PSIM_ERR_INVALID, "psim_generate_particles_device: bad array");
    if (num_parts == 0) return PSIM_OK;
    if (!pointer_on_device(parts_device)) return fail(PSIM_ERR_INVALID, "psim_generate_particles_device: needs a device array (psim_init_particles fills host arrays)");
    const int sx = (int)std::ceil(std::sqrt((double)num_parts));
    const int sy = (num_parts + sx - 1) / sx;
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < (unsigned long long)num_parts) ++half_bits;
    generate_particles_kernel<<<(num_parts + kObsThreads - 1) / kObsThreads, kObsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        parts_device, num_parts, size, (unsigned)seed, sx, sy, half_bits);
    PSIM_CUDA(cudaGetLastError());
    return PSIM_OK;
}

int psim_device_count(int* count) {
    if (!count) return fail(PSIM_ERR_INVALID, "psim_device_count: NULL");
    *count = 0;
    if (cudaGetDeviceCount(count) != cudaSuccess || *count == 0) {
        cudaGetLastError();
        return fail(PSIM_ERR_NO_DEVICE, "psim_device_count: no CUDA device visible; libpsim has no CPU fallback");
    }
    return PSIM_OK;
}

int psim_device_init(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PSIM_ERR_NO_DEVICE, "psim_device_init: no CUDA device visible; libpsim has no CPU fallback");
    }
    if (device >= ndev) return fail(PSIM_ERR_INVALID, "psim_device_init: device %d of %d", device, ndev);
    if (device >= 0) PSIM_CUDA(cudaSetDevice(device));
    PSIM_CUDA(cudaFree(nullptr));   // forces context creation
    return PSIM_OK;
}

int psim_create(psim_sim** out, const psim_config* cfg_in, const particle_t* parts, int num_parts, double size) {
    if (!out) return fail(PSIM_ERR_INVALID, "psim_create: out is NULL");
    *out = nullptr;
    psim_config cfg;
    if (cfg_in) cfg = *cfg_in;
    else psim_config_default(&cfg);
    if (cfg.nranks < 1) cfg.nranks = 1;
    if (num_parts < 0 || (num_parts > 0 && !parts)) return fail(PSIM_ERR_INVALID, "psim_create: bad particle array");
    if (!(size > 0.0) || !std::isfinite(size)) return fail(PSIM_ERR_INVALID, "psim_create: box size must be positive");
    if (cfg.rank < 0 || cfg.rank >= cfg.nranks) return fail(PSIM_ERR_INVALID, "psim_create: rank %d of %d", cfg.rank, cfg.nranks);
    const double nb = std::ceil(size / PSIM_BIN_SIZE);
    if (nb < 1 || nb * nb > 2.0e9) return fail(PSIM_ERR_UNSUPPORTED, "psim_create: %g cells per side do not fit int32 cell ids", nb);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PSIM_ERR_NO_DEVICE, "psim_create: no CUDA device visible; libpsim has no CPU fallback");
    }
    int dev = cfg.device;
    if (dev < 0) PSIM_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(PSIM_ERR_INVALID, "psim_create: device %d of %d", dev, ndev);
    PSIM_CUDA(cudaSetDevice(dev));

    psim_sim* sim = new psim_sim();
    sim->device = dev;
    sim->n_total = num_parts;
    sim->size = size;
    sim->bincnt = (int)nb;
    sim->rank = cfg.rank;
    sim->nranks = cfg.nranks;
    sim->row_begin = 0;
    sim->row_end = sim->bincnt;
    auto bail = [&](int st) {
        std::string keep = g_last_error;
        psim_destroy(sim);
        g_last_error = keep;
        return st;
    };
    if (cfg.stream) {
        sim->stream = static_cast<cudaStream_t>(cfg.stream);
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) return bail(fail(PSIM_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)));
        sim->own_stream = true;
    }
    int st = sim->mem.alloc(&sim->d_err, kErrWords);
    if (st) return bail(st);
    if (cudaMemsetAsync(sim->d_err, 0, kErrWords * sizeof(int), sim->stream) != cudaSuccess ||
        cudaHostAlloc(&sim->h_err, kErrWords * sizeof(int), cudaHostAllocDefault) != cudaSuccess)
        return bail(fail(PSIM_ERR_CUDA, "psim_create: error-word allocation failed: %s", cudaGetErrorString(cudaGetLastError())));

    const bool on_device = num_parts > 0 && pointer_on_device(parts);
    sim->input_on_device = on_device;
    int engine = cfg.engine;
    if (engine != PSIM_ENGINE_AUTO && engine != PSIM_ENGINE_CELLSORT && engine != PSIM_ENGINE_TILED && engine != PSIM_ENGINE_KSTEP)
        return bail(fail(PSIM_ERR_INVALID, "psim_create: unknown engine %d", engine));
    if (engine == PSIM_ENGINE_CELLSORT && cfg.nranks > 1)
        return bail(fail(PSIM_ERR_UNSUPPORTED, "psim_create: slabs need the kstep or the tiled engine"));
    if (engine == PSIM_ENGINE_AUTO) {
        if (const char* pref = std::getenv("PSIM_AUTO_ENGINE")) {   // A/B knob: what AUTO resolves to first
            if (!std::strcmp(pref, "tiled")) engine = PSIM_ENGINE_TILED;
            else if (!std::strcmp(pref, "cellsort") && cfg.nranks == 1) engine = PSIM_ENGINE_CELLSORT;
        }
    }

    if (engine == PSIM_ENGINE_AUTO || engine == PSIM_ENGINE_KSTEP) {
        bool unsuitable = false;
        st = kstep_create(sim, &cfg, parts, num_parts, on_device, &unsuitable);
        if (st == PSIM_OK) {
            sim->engine = PSIM_ENGINE_KSTEP;
        } else if (unsuitable && engine == PSIM_ENGINE_AUTO && cfg.nranks == 1) {
            kstep_destroy(sim);
            engine = PSIM_ENGINE_CELLSORT;
        } else {
            return bail(st);
        }
    }
    if (engine == PSIM_ENGINE_TILED) {
        bool unsuitable = false;
        st = tiled_create(sim, &cfg, parts, num_parts, on_device, &unsuitable);
        if (st == PSIM_OK) {
            sim->engine = PSIM_ENGINE_TILED;
        } else {
            return bail(st);
        }
    }
    if (engine == PSIM_ENGINE_CELLSORT) {
        const particle_t* d_parts = parts;
        DeviceArena stage;
        if (!on_device && num_parts > 0) {
            particle_t* tmp = nullptr;
            st = stage.alloc(&tmp, (size_t)num_parts);
            if (st) return bail(st);
            cudaError_t e = cudaMemcpyAsync(tmp, parts, sizeof(particle_t) * (size_t)num_parts, cudaMemcpyHostToDevice, sim->stream);
            if (e != cudaSuccess) {
                stage.release();
                return bail(fail(PSIM_ERR_CUDA, "upload: %s", cudaGetErrorString(e)));
            }
            d_parts = tmp;
        }
        st = cellsort_create(sim, d_parts, num_parts);
        cudaStreamSynchronize(sim->stream);
        stage.release();
        if (st) return bail(st);
        sim->engine = PSIM_ENGINE_CELLSORT;
    }
    st = check_device_error(sim);
    if (st) return bail(st);
    *out = sim;
    return PSIM_OK;
}

int psim_destroy(psim_sim* sim) {
    if (!sim) return PSIM_OK;
    cudaSetDevice(sim->device);
    if (sim->stream) cudaStreamSynchronize(sim->stream);
    comm_destroy(sim);
    cellsort_destroy(sim);
    tiled_destroy(sim);
    kstep_destroy(sim);
    for (int k = 0; k < 2; ++k) {
        if (sim->async_ready[k]) cudaEventDestroy(sim->async_ready[k]);
        if (sim->async_done[k]) cudaEventDestroy(sim->async_done[k]);
    }
    if (sim->copy_stream) {
        cudaStreamSynchronize(sim->copy_stream);
        cudaStreamDestroy(sim->copy_stream);
    }
    if (sim->async_err) cudaFreeHost(sim->async_err);
    sim->scratch.release();
    sim->mem.release();
    if (sim->h_err) cudaFreeHost(sim->h_err);
    if (sim->own_stream && sim->stream) cudaStreamDestroy(sim->stream);
    delete sim;
    return PSIM_OK;
}

int psim_step(psim_sim* sim, int nsteps, int flags) {
    if (!sim) return fail(PSIM_ERR_INVALID, "psim_step: NULL handle");
    if (nsteps < 0) return fail(PSIM_ERR_INVALID, "psim_step: nsteps %d", nsteps);
    if (nsteps == 0) return PSIM_OK;
    DeviceGuard g(sim->device);
    if (sim->engine == PSIM_ENGINE_KSTEP) return kstep_step(sim, nsteps, flags);
    return sim->engine == PSIM_ENGINE_CELLSORT ? cellsort_step(sim, nsteps, flags) : tiled_step(sim, nsteps, flags);
}

int psim_sync(psim_sim* sim) {
    if (!sim) return fail(PSIM_ERR_INVALID, "psim_sync: NULL handle");
    DeviceGuard g(sim->device);
    return check_device_error(sim);
}

int psim_read_particles(psim_sim* sim, particle_t* dst) {
    if (!sim || !dst) return fail(PSIM_ERR_INVALID, "psim_read_particles: NULL argument");
    DeviceGuard g(sim->device);
    PSIM_TRY(check_device_error(sim));
    if ((sim->engine == PSIM_ENGINE_TILED || sim->engine == PSIM_ENGINE_KSTEP) && sim->nranks == 1) {
        // one pass straight from the tiles into the caller's order (no intermediate compaction)
        cudaStream_t s = sim->stream;
        if (pointer_on_device(dst)) {
            PSIM_TRY(engine_writeback(sim, dst, nullptr));
            PSIM_CUDA(cudaStreamSynchronize(s));
            return PSIM_OK;
        }
        particle_t* stage = nullptr;
        PSIM_TRY(ensure_scratch(sim, sizeof(particle_t) * (size_t)std::max(sim->n_total, 1), reinterpret_cast<void**>(&stage)));
        PSIM_TRY(engine_writeback(sim, stage, nullptr));
        PSIM_CUDA(cudaMemcpyAsync(dst, stage, sizeof(particle_t) * (size_t)sim->n_total, cudaMemcpyDeviceToHost, s));
        PSIM_CUDA(cudaStreamSynchronize(s));
        return PSIM_OK;
    }
    if ((sim->engine == PSIM_ENGINE_TILED || sim->engine == PSIM_ENGINE_KSTEP) && sim->nranks > 1) {
        // a slab writes only the records it owns (the analogue of the reference's gather_for_save, part2/mpi.cpp:371-402):
        // straight from the stripes into the caller's array, in original order -- a device array, or a pinned host array
        // through its device alias (zero-copy scatter over PCIe: no staging buffer, no host loop)
        particle_t* d = pointer_on_device(dst) ? dst : static_cast<particle_t*>(device_alias_of_host(dst));
        if (d) {
            PSIM_TRY(engine_writeback(sim, d, nullptr));
            PSIM_CUDA(cudaStreamSynchronize(sim->stream));
            return PSIM_OK;
        }
    }
    SoAView v;
    bool acc;
    PSIM_TRY(view_of(sim, &v, &acc));
    cudaStream_t s = sim->stream;
    const int blocks = (v.n + kObsThreads - 1) / kObsThreads;
    if (pointer_on_device(dst)) {
        if (v.n) soa_to_aos_kernel<<<blocks, kObsThreads, 0, s>>>(v, acc, dst);
        ++sim->launches;
        PSIM_CUDA(cudaGetLastError());
        PSIM_CUDA(cudaStreamSynchronize(s));
        return PSIM_OK;
    }
    particle_t* stage = nullptr;
    if (sim->nranks == 1) {
        PSIM_TRY(ensure_scratch(sim, sizeof(particle_t) * (size_t)std::max(sim->n_total, 1), reinterpret_cast<void**>(&stage)));
        if (v.n) soa_to_aos_kernel<<<blocks, kObsThreads, 0, s>>>(v, acc, stage);
        ++sim->launches;
        PSIM_CUDA(cudaGetLastError());
        PSIM_CUDA(cudaMemcpyAsync(dst, stage, sizeof(particle_t) * (size_t)sim->n_total, cudaMemcpyDeviceToHost, s));
        PSIM_CUDA(cudaStreamSynchronize(s));
    } else {
        // a slab writes only the records it owns: pack on the device, place on the host
        PSIM_TRY(ensure_scratch(sim, sizeof(particle_t) * (size_t)std::max(v.n, 1), reinterpret_cast<void**>(&stage)));
        if (v.n) soa_pack_kernel<<<blocks, kObsThreads, 0, s>>>(v, acc, stage);
        ++sim->launches;
        // pageable destination: pinned staging for the packed records and their ids, then the host places them
        particle_t* h_rec = nullptr;
        int* h_ids = nullptr;
        PSIM_CUDA(cudaMallocHost(&h_rec, sizeof(particle_t) * (size_t)std::max(v.n, 1)));
        if (cudaMallocHost(&h_ids, sizeof(int) * (size_t)std::max(v.n, 1)) != cudaSuccess) {
            cudaFreeHost(h_rec);
            return fail(PSIM_ERR_CUDA, "psim_read_particles: pinned staging: %s", cudaGetErrorString(cudaGetLastError()));
        }
        cudaError_t ce = cudaMemcpyAsync(h_rec, stage, sizeof(particle_t) * (size_t)v.n, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_ids, v.id, sizeof(int) * (size_t)v.n, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        if (ce == cudaSuccess) {
            const int nthreads = std::max(1, std::min(8, v.n / (1 << 18)));
            std::vector<std::thread> pool;
            for (int w = 0; w < nthreads; ++w)
                pool.emplace_back([=]() {
                    const long long lo = (long long)v.n * w / nthreads, hi = (long long)v.n * (w + 1) / nthreads;
                    for (long long i = lo; i < hi; ++i) dst[h_ids[i]] = h_rec[i];
                });
            for (auto& t : pool) t.join();
        }
        cudaFreeHost(h_rec);
        cudaFreeHost(h_ids);
        if (ce != cudaSuccess) return fail(PSIM_ERR_CUDA, "psim_read_particles: %s", cudaGetErrorString(ce));
    }
    return PSIM_OK;
}

int psim_read_positions(psim_sim* sim, double* xy) {
    if (!sim || !xy) return fail(PSIM_ERR_INVALID, "psim_read_positions: NULL argument");
    DeviceGuard g(sim->device);
    PSIM_TRY(check_device_error(sim));
    if ((sim->engine == PSIM_ENGINE_TILED || sim->engine == PSIM_ENGINE_KSTEP) && sim->nranks == 1) {
        cudaStream_t s = sim->stream;
        if (pointer_on_device(xy)) {
            PSIM_TRY(engine_writeback(sim, nullptr, reinterpret_cast<double2*>(xy)));
            PSIM_CUDA(cudaStreamSynchronize(s));
            return PSIM_OK;
        }
        double2* stage = nullptr;
        PSIM_TRY(ensure_scratch(sim, sizeof(double2) * (size_t)std::max(sim->n_total, 1), reinterpret_cast<void**>(&stage)));
        PSIM_TRY(engine_writeback(sim, nullptr, stage));
        PSIM_CUDA(cudaMemcpyAsync(xy, stage, sizeof(double2) * (size_t)sim->n_total, cudaMemcpyDeviceToHost, s));
        PSIM_CUDA(cudaStreamSynchronize(s));
        return PSIM_OK;
    }
    if ((sim->engine == PSIM_ENGINE_TILED || sim->engine == PSIM_ENGINE_KSTEP) && sim->nranks > 1) {
        double2* d = pointer_on_device(xy) ? reinterpret_cast<double2*>(xy) : static_cast<double2*>(device_alias_of_host(xy));
        if (d) {
            PSIM_TRY(engine_writeback(sim, nullptr, d));
            PSIM_CUDA(cudaStreamSynchronize(sim->stream));
            return PSIM_OK;
        }
    }
    SoAView v;
    bool acc;
    PSIM_TRY(view_of(sim, &v, &acc));
    cudaStream_t s = sim->stream;
    const int blocks = (v.n + kObsThreads - 1) / kObsThreads;
    if (pointer_on_device(xy)) {
        if (v.n) soa_to_xy_kernel<<<blocks, kObsThreads, 0, s>>>(v, reinterpret_cast<double2*>(xy));
        ++sim->launches;
        PSIM_CUDA(cudaGetLastError());
        PSIM_CUDA(cudaStreamSynchronize(s));
        return PSIM_OK;
    }
    double2* stage = nullptr;
    PSIM_TRY(ensure_scratch(sim, sizeof(double2) * (size_t)std::max(sim->n_total, 1), reinterpret_cast<void**>(&stage)));
    if (sim->nranks > 1)
        PSIM_CUDA(cudaMemcpyAsync(stage, xy, sizeof(double2) * (size_t)sim->n_total, cudaMemcpyHostToDevice, s));
    if (v.n) soa_to_xy_kernel<<<blocks, kObsThreads, 0, s>>>(v, stage);
    ++sim->launches;
    PSIM_CUDA(cudaGetLastError());
    PSIM_CUDA(cudaMemcpyAsync(xy, stage, sizeof(double2) * (size_t)sim->n_total, cudaMemcpyDeviceToHost, s));
    PSIM_CUDA(cudaStreamSynchronize(s));
    return PSIM_OK;
}

// Asynchronous flavour of psim_read_positions for the save path (reference part3/main.cu:134-136 copies the whole array
// synchronously before every save): _begin enqueues the original-order gather on the handle's stream and the device->host copy
// on a separate copy stream and returns; steps enqueued afterwards overlap the copy.  _end waits for the OLDEST read that has
// not been waited for.  At most two reads may be in flight; `xy_host` should be page-locked (psim_host_register) or the copy
// degrades to a staged, synchronous one inside the CUDA runtime.
int psim_read_positions_begin(psim_sim* sim, double* xy_host) {
    if (!sim || !xy_host) return fail(PSIM_ERR_INVALID, "psim_read_positions_begin: NULL argument");
    DeviceGuard g(sim->device);
    if (sim->async_issued - sim->async_waited >= 2) return fail(PSIM_ERR_STATE, "psim_read_positions_begin: two reads already in flight");
    if (sim->nranks > 1 || sim->engine == PSIM_ENGINE_CELLSORT) {   // (slabs / the fallback engine: plain synchronous read)
        PSIM_TRY(psim_read_positions(sim, xy_host));
        ++sim->async_issued;
        return PSIM_OK;
    }
    if (!sim->copy_stream) {
        PSIM_CUDA(cudaStreamCreateWithFlags(&sim->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            PSIM_TRY(sim->mem.alloc(&sim->async_stage[k], (size_t)std::max(sim->n_total, 1)));
            PSIM_CUDA(cudaEventCreateWithFlags(&sim->async_ready[k], cudaEventDisableTiming));
            PSIM_CUDA(cudaEventCreateWithFlags(&sim->async_done[k], cudaEventDisableTiming));
        }
        PSIM_CUDA(cudaHostAlloc(&sim->async_err, 2 * kErrWords * sizeof(int), cudaHostAllocDefault));
    }
    const int k = (int)(sim->async_issued & 1u);
    // (the copy that used this staging buffer two reads ago has been waited for: in-flight count < 2)
    PSIM_TRY(engine_writeback(sim, nullptr, sim->async_stage[k]));
    PSIM_CUDA(cudaEventRecord(sim->async_ready[k], sim->stream));
    PSIM_CUDA(cudaStreamWaitEvent(sim->copy_stream, sim->async_ready[k], 0));
    PSIM_CUDA(cudaMemcpyAsync(xy_host, sim->async_stage[k], sizeof(double2) * (size_t)sim->n_total, cudaMemcpyDeviceToHost, sim->copy_stream));
    // the device error words as of this read travel with it: a launch that hit a bound leaves a state that must not be saved
    PSIM_CUDA(cudaMemcpyAsync(sim->async_err + k * kErrWords, sim->d_err, kErrWords * sizeof(int), cudaMemcpyDeviceToHost, sim->copy_stream));
    PSIM_CUDA(cudaEventRecord(sim->async_done[k], sim->copy_stream));
    sim->async_dst[k] = xy_host;
    ++sim->async_issued;
    return PSIM_OK;
}

int psim_read_positions_end(psim_sim* sim) {
    if (!sim) return fail(PSIM_ERR_INVALID, "psim_read_positions_end: NULL handle");
    if (sim->async_issued == sim->async_waited) return fail(PSIM_ERR_STATE, "psim_read_positions_end: no read in flight");
    DeviceGuard g(sim->device);
    const int k = (int)(sim->async_waited & 1u);
    ++sim->async_waited;
    if (!sim->async_dst[k]) return PSIM_OK;   // that read was done synchronously
    PSIM_CUDA(cudaEventSynchronize(sim->async_done[k]));
    double* dst = sim->async_dst[k];
    sim->async_dst[k] = nullptr;
    const int* e = sim->async_err + k * kErrWords;
    if (e[0] != 0 || e[8] != 0) {
        // a launch before this read hit a capacity / speed bound: let the synchronous path replay it (or hand the state over to
        // the cellsort engine) and read the repaired state.  NOTE: this also includes the steps enqueued after _begin; callers
        // that need the exact frame should treat this rare event as "frame taken late".
        PSIM_TRY(psim_read_positions(sim, dst));
    }
    return PSIM_OK;
}

int psim_read_cells(psim_sim* sim, int* cell_of_particle, int* cell_counts) {
    if (!sim) return fail(PSIM_ERR_INVALID, "psim_read_cells: NULL handle");
    DeviceGuard g(sim->device);
    PSIM_TRY(check_device_error(sim));
    SoAView v;
    bool acc;
    PSIM_TRY(view_of(sim, &v, &acc));
    cudaStream_t s = sim->stream;
    const int blocks = (v.n + kObsThreads - 1) / kObsThreads;
    if (cell_of_particle) {
        DeviceArena tmp;
        int* d = nullptr;
        PSIM_TRY(tmp.alloc(&d, (size_t)sim->n_total));
        PSIM_CUDA(cudaMemsetAsync(d, 0xFF, sizeof(int) * (size_t)sim->n_total, s));  // -1 = not owned by this slab
        if (v.n) cell_of_particle_kernel<<<blocks, kObsThreads, 0, s>>>(v, sim->bincnt, d);
        ++sim->launches;
        PSIM_CUDA(cudaGetLastError());
        PSIM_CUDA(cudaMemcpyAsync(cell_of_particle, d, sizeof(int) * (size_t)sim->n_total, cudaMemcpyDeviceToHost, s));
        PSIM_CUDA(cudaStreamSynchronize(s));
        tmp.release();
    }
    if (cell_counts) {
        CellBinner b;
        int st = b.init(sim->bincnt, std::max(v.n, 1));
        if (st == PSIM_OK) st = b.count(v.x, v.y, v.n, s);
        if (st == PSIM_OK) {
            cudaError_t e = cudaMemcpyAsync(cell_counts, b.cell_start, sizeof(int) * (size_t)b.ncell, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) st = fail(PSIM_ERR_CUDA, "psim_read_cells: %s", cudaGetErrorString(e));
        }
        sim->launches += b.launches;
        b.release();
        PSIM_TRY(st);
    }
    return PSIM_OK;
}

int psim_read_cell_lists(psim_sim* sim, int* cell_start, int* members) {
    if (!sim || !cell_start || !members) return fail(PSIM_ERR_INVALID, "psim_read_cell_lists: NULL argument");
    DeviceGuard g(sim->device);
    PSIM_TRY(check_device_error(sim));
    SoAView v;
    bool acc;
    PSIM_TRY(view_of(sim, &v, &acc));
    cudaStream_t s = sim->stream;
    const int blocks = (v.n + kObsThreads - 1) / kObsThreads;
    CellBinner b;
    DeviceArena tmp;
    int *raw = nullptr, *mem = nullptr;
    int st = b.init(sim->bincnt, std::max(v.n, 1));
    if (st == PSIM_OK) st = b.build(v.x, v.y, v.n, s);
    if (st == PSIM_OK) st = tmp.alloc(&raw, (size_t)v.n);
    if (st == PSIM_OK) st = tmp.alloc(&mem, (size_t)v.n);
    if (st == PSIM_OK && v.n) {
        members_raw_kernel<<<blocks, kObsThreads, 0, s>>>(v, sim->bincnt, b.cell_start, b.slot, raw);
        members_rank_kernel<<<blocks, kObsThreads, 0, s>>>(v, sim->bincnt, b.cell_start, raw, mem);
        sim->launches += 2;
    }
    if (st == PSIM_OK) {
        cudaError_t e = cudaMemcpyAsync(cell_start, b.cell_start, sizeof(int) * ((size_t)b.ncell + 1), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && v.n) e = cudaMemcpyAsync(members, mem, sizeof(int) * (size_t)v.n, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) st = fail(PSIM_ERR_CUDA, "psim_read_cell_lists: %s", cudaGetErrorString(e));
    }
    sim->launches += b.launches;
    b.release();
    tmp.release();
    return st;
}

int psim_stats(psim_sim* sim, psim_stats_t* out) {
    if (!sim || !out) return fail(PSIM_ERR_INVALID, "psim_stats: NULL argument");
    DeviceGuard g(sim->device);
    PSIM_TRY(check_device_error(sim));
    SoAView v;
    bool acc;
    PSIM_TRY(view_of(sim, &v, &acc));
    cudaStream_t s = sim->stream;
    std::memset(out, 0, sizeof *out);
    out->dmin = 1.0;
    if (v.n == 0) return PSIM_OK;
    CellBinner b;
    DeviceArena tmp;
    double *sx = nullptr, *sy = nullptr;
    StatsPartial* part = nullptr;
    int st = PSIM_OK;
    // A slab of the kstep engine also looks at the particles in its ghost rows (the neighbours' boundary bands), as
    // neighbours only: pairs that straddle a slab border are then counted on both sides, like inside one slab, and the
    // per-slab numbers of all ranks add up to the single-GPU numbers.
    const int n_owned = v.n;
    SoAView all = v;
    unsigned char* owned_sorted = nullptr;
    if (sim->engine == PSIM_ENGINE_KSTEP && sim->nranks > 1) {
        const int ghost_cap = kstep_ghost_capacity(sim);
        double *cx = nullptr, *cy = nullptr;
        int n_ghost = 0;
        st = tmp.alloc(&cx, (size_t)n_owned + ghost_cap);
        if (st == PSIM_OK) st = tmp.alloc(&cy, (size_t)n_owned + ghost_cap);
        if (st == PSIM_OK) st = tmp.alloc(&owned_sorted, (size_t)n_owned + ghost_cap);
        if (st == PSIM_OK && (cudaMemcpyAsync(cx, v.x, sizeof(double) * (size_t)n_owned, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
                              cudaMemcpyAsync(cy, v.y, sizeof(double) * (size_t)n_owned, cudaMemcpyDeviceToDevice, s) != cudaSuccess))
            st = fail(PSIM_ERR_CUDA, "psim_stats: %s", cudaGetErrorString(cudaGetLastError()));
        if (st == PSIM_OK) st = kstep_ghost_positions(sim, cx + n_owned, cy + n_owned, ghost_cap, &n_ghost);
        if (st != PSIM_OK) {
            tmp.release();
            return st;
        }
        all.x = cx;
        all.y = cy;
        all.n = n_owned + n_ghost;
    }
    const int blocks = (all.n + kObsThreads - 1) / kObsThreads;
    if (st == PSIM_OK) st = b.init(sim->bincnt, all.n);
    if (st == PSIM_OK) st = b.build(all.x, all.y, all.n, s);
    if (st == PSIM_OK) st = tmp.alloc(&sx, (size_t)all.n);
    if (st == PSIM_OK) st = tmp.alloc(&sy, (size_t)all.n);
    if (st == PSIM_OK) st = tmp.alloc(&part, (size_t)blocks);
    std::vector<StatsPartial> h((size_t)blocks);
    if (st == PSIM_OK) {
        sort_xy_kernel<<<blocks, kObsThreads, 0, s>>>(all, sim->bincnt, b.cell_start, b.slot, sx, sy, owned_sorted, n_owned);
        stats_kernel<<<blocks, kObsThreads, 0, s>>>(sx, sy, all.n, v, sim->bincnt, b.cell_start, owned_sorted, n_owned, part);
        sim->launches += 2;
        cudaError_t e = cudaMemcpyAsync(h.data(), part, sizeof(StatsPartial) * (size_t)blocks, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) st = fail(PSIM_ERR_CUDA, "psim_stats: %s", cudaGetErrorString(e));
    }
    sim->launches += b.launches;
    b.release();
    tmp.release();
    PSIM_TRY(st);
    double dsum = 0;
    for (const StatsPartial& p : h) {
        out->dmin = std::min(out->dmin, p.dmin);
        dsum += p.dsum;
        out->kinetic_energy += p.ke;
        out->vmax = std::max(out->vmax, p.vmax);
        out->pairs += p.pairs;
        out->touched += p.touched;
        out->max_neighbours = std::max(out->max_neighbours, p.maxnb);
        out->max_cell_count = std::max(out->max_cell_count, p.maxcell);
    }
    out->davg = out->pairs ? dsum / (double)out->pairs : 0.0;
    return PSIM_OK;
}

int psim_gather(psim_sim* sim, particle_t* dst, int root) {
    if (!sim) return fail(PSIM_ERR_INVALID, "psim_gather: NULL handle");
    if (root < 0 || root >= sim->nranks) return fail(PSIM_ERR_INVALID, "psim_gather: root %d of %d", root, sim->nranks);
    if (sim->rank == root && !dst) return fail(PSIM_ERR_INVALID, "psim_gather: the root needs a destination");
    if (sim->nranks == 1) return psim_read_particles(sim, dst);
    DeviceGuard g(sim->device);
    PSIM_TRY(check_device_error(sim));
    SoAView v;
    bool acc;
    PSIM_TRY(view_of(sim, &v, &acc));
    cudaStream_t s = sim->stream;
    std::vector<int> counts((size_t)sim->nranks);
    PSIM_TRY(comm_allgather_counts(sim, v.n, counts.data(), s));
    long long total = 0;
    for (int c : counts) total += c;
    if (total != sim->n_total) return fail(PSIM_ERR_STATE, "psim_gather: the slabs own %lld particles, expected %d", total, sim->n_total);
    const bool is_root = sim->rank == root;
    DeviceArena tmp;
    particle_t *rec = nullptr, *out = nullptr;
    int* ids = nullptr;
    // root: room for everybody's records (its own block at its offset); others: their own block
    const size_t cap = is_root ? (size_t)sim->n_total : (size_t)std::max(v.n, 1);
    int st = tmp.alloc(&rec, cap);
    if (st == PSIM_OK) st = tmp.alloc(&ids, cap);
    size_t my_off = 0;
    if (is_root)
        for (int r = 0; r < root; ++r) my_off += (size_t)counts[r];
    if (st == PSIM_OK && v.n) {
        soa_pack_kernel<<<(v.n + kObsThreads - 1) / kObsThreads, kObsThreads, 0, s>>>(v, acc, rec + my_off);
        ++sim->launches;
        if (cudaMemcpyAsync(ids + my_off, v.id, sizeof(int) * (size_t)v.n, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
            st = fail(PSIM_ERR_CUDA, "psim_gather: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (st == PSIM_OK) st = comm_gather_records(sim, root, rec, ids, v.n, sizeof(particle_t), rec, ids, counts.data(), s);
    if (st == PSIM_OK && is_root) {
        const bool on_dev = pointer_on_device(dst);
        out = dst;
        if (!on_dev) st = tmp.alloc(&out, (size_t)sim->n_total);
        if (st == PSIM_OK) {
            const long long threads = 3ll * sim->n_total;
            place_records_kernel<<<(unsigned)((threads + kObsThreads - 1) / kObsThreads), kObsThreads, 0, s>>>(rec, ids, sim->n_total, out);
            ++sim->launches;
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess && !on_dev) e = cudaMemcpyAsync(dst, out, sizeof(particle_t) * (size_t)sim->n_total, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) st = fail(PSIM_ERR_CUDA, "psim_gather: %s", cudaGetErrorString(e));
        }
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && st == PSIM_OK) st = fail(PSIM_ERR_CUDA, "psim_gather: %s", cudaGetErrorString(cudaGetLastError()));
    tmp.release();
    return st;
}

int psim_state_hash(psim_sim* sim, unsigned long long* out, long long* owned) {
    if (!sim || !out) return fail(PSIM_ERR_INVALID, "psim_state_hash: NULL argument");
    DeviceGuard g(sim->device);
    PSIM_TRY(check_device_error(sim));
    SoAView v;
    bool acc;
    PSIM_TRY(view_of(sim, &v, &acc));
    cudaStream_t s = sim->stream;
    DeviceArena tmp;
    unsigned long long* d = nullptr;
    PSIM_TRY(tmp.alloc(&d, 1));
    int st = PSIM_OK;
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long), s);
    if (e == cudaSuccess && v.n) {
        state_hash_kernel<<<(v.n + kObsThreads - 1) / kObsThreads, kObsThreads, 0, s>>>(v, d);
        ++sim->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) st = fail(PSIM_ERR_CUDA, "psim_state_hash: %s", cudaGetErrorString(e));
    tmp.release();
    if (owned) *owned = v.n;
    return st;
}

int psim_info(psim_sim* sim, psim_info_t* out) {
    if (!sim || !out) return fail(PSIM_ERR_INVALID, "psim_info: NULL argument");
    std::memset(out, 0, sizeof *out);
    out->engine = sim->engine;
    out->bin_count = sim->bincnt;
    out->device = sim->device;
    out->rank = sim->rank;
    out->nranks = sim->nranks;
    out->row_begin = sim->row_begin;
    out->row_end = sim->row_end;
    out->steps_done = sim->steps_done;
    out->kernel_launches = sim->launches;
    out->num_parts = sim->nranks == 1 ? sim->n_total : sim->owned_last;   // slabs: as of the last observation call
    out->engine_switches = sim->engine_switches;
    out->input_on_device = sim->input_on_device ? 1 : 0;
    out->device_bytes = (long long)sim->mem.bytes + cellsort_bytes(sim) + tiled_bytes(sim) + kstep_bytes(sim);
    if (sim->engine == PSIM_ENGINE_KSTEP) {
        kstep_info(sim, out);
        out->hw_tile_population = sim->h_err[3];
        out->hw_apron = sim->h_err[4];
        out->reserved_hw_pairs = sim->h_err[7];
    }
    if (sim->engine == PSIM_ENGINE_TILED) {
        tiled_info(sim, out);
        out->hw_leavers = sim->h_err[1];
        out->hw_halo_list = sim->h_err[2];
        out->hw_tile_population = sim->h_err[3];
        out->hw_apron = sim->h_err[4];
        out->reserved_hw_pairs = sim->h_err[7];
    }
    return PSIM_OK;
}

}  // extern "C"
