/*
 * wc_sph.h -- C-ABI of the B200-native WaterCube SPH step.
 *
 * The reference has no plugin/FFI ABI: its boundary is the C++ class surface of
 * core::Fluid (src/core/Fluid.h:32-59) and core::Sort (src/core/Sort.h:20-39), which
 * drive five GLSL compute dispatches per frame (src/core/Fluid.cpp:342-354,
 * src/core/Sort.cpp:254-267).  This header is the drop-in boundary for that path:
 * plain pointers and sizes, an opaque handle, `int` status (0 = ok) plus
 * wc_last_error().  The C++ facade (watercube_b200/csrc/core/Fluid.h, Sort.h, util.h)
 * keeps the reference's method names on top of exactly these calls; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Threading: a handle is not thread-safe (the reference calls Fluid::update from the
 * single Cinder main thread, src/WaterCubeApp.cpp:98).  All work of a handle is ordered
 * on one CUDA stream.  There is NO CPU fallback: every call fails with WC_ERR_NO_DEVICE /
 * WC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef WC_SPH_H
#define WC_SPH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WC_ABI_VERSION 1

enum {
    WC_OK = 0,
    WC_ERR_INVALID = 1,   /* bad argument / call order */
    WC_ERR_NO_DEVICE = 2, /* no CUDA device, or device ordinal out of range */
    WC_ERR_CUDA = 3,      /* a CUDA runtime call or kernel failed; see wc_last_error() */
    WC_ERR_CAPACITY = 4   /* particle count exceeds the handle's capacity */
};

/* wc_params.flags */
enum {
    WC_FLAG_DEBUG_OUTPUTS = 1, /* also store neighbour counts and total forces (parity runs) */
    WC_FLAG_STAGE_TIMING = 2,  /* record CUDA events around every stage of wc_step */
    WC_FLAG_SIMPLE_KERNELS = 4 /* use the simple thread-per-particle gather kernels
                                  (cross-check for the warp-cooperative tiled ones) */
};

/* Stages of one step, in execution order (index into wc_stage_times). */
enum {
    WC_STAGE_HASH_COUNT = 0, /* count.comp (+ counter clear, Sort.cpp:255-256) */
    WC_STAGE_SCAN = 1,       /* linearScan.comp (Sort.cpp:258-259) */
    WC_STAGE_REORDER = 2,    /* reorder.comp, made stable (Sort.cpp:263-264) */
    WC_STAGE_DENSITY = 3,    /* density.comp (Fluid.cpp:349) */
    WC_STAGE_UPDATE = 4,     /* update.comp (Fluid.cpp:350) */
    WC_NUM_STAGES = 5
};

/* 32-byte AoS particle: struct Particle, src/core/util.h:29-35 (GLSL mirror
 * assets/fluid/density.comp:5-10).  This is the host-visible particle-buffer layout of
 * util::getParticles / util::setParticles (src/core/util.cpp:42-63). */
typedef struct wc_particle {
    float position[3];
    float density;
    float velocity[3];
    float pressure;
} wc_particle;

/* Setup-time parameters: Fluid's constructor defaults and fluent setters
 * (src/core/Fluid.cpp:9-27, :31-84).  N, gridRes, size and radius are fixed at setup
 * (Fluid::setup, Fluid.cpp:203-235). */
typedef struct wc_params {
    int32_t num_particles;  /* Fluid::numParticles(int), default 80000 */
    int32_t capacity;       /* particle slots to allocate; 0 = num_particles */
    int32_t grid_res;       /* Fluid::gridRes(int), default 21 */
    float size;             /* Fluid::size(float), default 1 */
    float particle_radius;  /* Fluid::particleRadius(float), default 0.01 */
    float time_scale;       /* Fluid.cpp:24, default 0.012: dt = frame_dt * time_scale */
    int32_t device;         /* CUDA device ordinal */
    uint32_t flags;         /* WC_FLAG_* */
    int32_t neighbour_list_words; /* per 32-particle group: capacity of the density->update
                               neighbour list in 32-candidate words; 0 = default (32..96,
                               from the scene's mean number density), < 0 = no list
                               (update repeats the search) */
    /* z-slab mode (multi-GPU, one handle per rank): this handle owns the global z-layers
     * [slab_z_begin, slab_z_end) of the grid.  Enabled when slab_ghost_capacity > 0;
     * `capacity` then bounds the OWNED particles.  See the wc_slab_* calls below. */
    int32_t slab_z_begin, slab_z_end;
    int32_t slab_ghost_capacity;   /* max particles of ONE neighbouring halo layer */
    int32_t slab_migrant_capacity; /* max particles crossing ONE slab face per step */
    void* stream;           /* optional caller-owned cudaStream_t; NULL = library-owned */
} wc_params;

/* Per-step parameters: the fields the reference re-uploads as uniforms on every
 * dispatch (GUI-mutable, src/core/Fluid.cpp:89-99, :276-285, :305-317). */
typedef struct wc_step_params {
    float viscosity_coefficient; /* default 200 */
    float stiffness;             /* default 100 */
    float rest_density;          /* default 500 */
    float rest_pressure;         /* default 0 */
    float gravity[3];            /* gravity_direction_ * gravity_strength_, default (0,-900,0) */
    float mouse_origin[3];       /* "cameraPosition" uniform: mouse ray origin, box space */
    float mouse_dir[3];          /* "mouseRayDirection" uniform */
} wc_step_params;

/* Physics the reference's report lists as future work (report.pdf section 6, "Future Work":
 * wall particles after Harada et al., surface tension after Yan et al.).
 * There is no reference code for it: oracle/wc_oracle.h defines the arithmetic, the CUDA path
 * is checked against that.  A handle starts with flags == 0, which is the reference's step bit
 * for bit (its own kernel instantiations, untouched by any of this); wc_set_physics switches
 * the handle to the instantiations that know the flags.
 *
 * WC_PHYS_WALL_PARTICLES replaces the pseudo wall density of density.comp:57-79 and the wall
 * force of update.comp:71-100 (quirks Q2-Q6): the wall particles' density contribution is a
 * function of the distance s to the wall only (Harada's "wall weight function", here the
 * closed-form integral of wall_rest_density * W_poly6 over the half space behind the wall),
 * and a particle closer than wall_distance is pushed back with the acceleration
 * wall_stiffness * (wall_distance - s) / dt^2 along the wall normal.  The box clamp of
 * update.comp:202-227 stays as the last resort.
 *
 * WC_PHYS_SURFACE_TENSION adds the colour-field force of Mueller et al. 2003 (the model Yan et
 * al. build on): n_i = sum_j m/rho_j grad W_poly6(r_ij), lap_i = sum_j m/rho_j lap W_poly6(r_ij)
 * over the particle itself and its neighbours with 0 < |r_ij| < h (summed like the density:
 * self term included), and F_i += -surface_tension * lap_i * n_i / |n_i| where
 * |n_i| > surface_threshold. */
#define WC_PHYS_WALL_PARTICLES 1u
#define WC_PHYS_SURFACE_TENSION 2u
typedef struct wc_physics {
    uint32_t flags;          /* WC_PHYS_*; 0 = the reference's physics */
    float surface_tension;   /* default 50 */
    float surface_threshold; /* default 7 */
    float wall_stiffness;    /* default 0.5: half of the penetration undone per step */
    float wall_distance;     /* default 0.01 (the reference's particle radius) */
    float wall_rest_density; /* mass x number density of the wall particles; <= 0: rest_density */
} wc_physics;

/* Constants derived in Fluid::setup (src/core/Fluid.cpp:206-216). */
typedef struct wc_derived {
    int32_t num_bins;
    float bin_size;
    float kernel_radius;
    float particle_mass;
    float poly6_const;
    float spiky_const;
    float visc_const;
    float dist2_threshold; /* smallest fp32 x with sqrtf(x) >= kernel_radius */
} wc_derived;

/* Raw device pointers for zero-copy consumers (future CUDA-GL interop of
 * Fluid::renderParticles, Fluid.cpp:389-406).  SoA float4: pos_rho = (x,y,z,density),
 * vel_pres = (vx,vy,vz,pressure).  Index 0 = "buffer 1" (current state), 1 = "buffer 2"
 * (cell-sorted input of the last step with density/pressure filled in). */
typedef struct wc_device_view {
    void* pos_rho[2];
    void* vel_pres[2];
    void* cell_ids;         /* uint32[num_particles], per INPUT particle of the last sort */
    void* counts;           /* uint32[num_bins]      Sort::getCountBuffer  */
    void* offsets;          /* uint32[num_bins + 1]  Sort::getOffsetBuffer (+ total sentinel) */
    void* sorted;           /* uint32[num_particles] Sort::getSortedBuffer: sorted[dst] = src */
    void* neighbour_counts; /* uint32[num_particles] (WC_FLAG_DEBUG_OUTPUTS) or NULL */
    void* forces;           /* float4[num_particles] (WC_FLAG_DEBUG_OUTPUTS) or NULL */
    void* stream;           /* the cudaStream_t all work is ordered on */
    int32_t num_particles;
    int32_t capacity;
} wc_device_view;

int wc_abi_version(void);
const char* wc_last_error(void);

/* Fluid::Fluid defaults (Fluid.cpp:9-27); mouse ray defaults to one that misses the box. */
int wc_default_params(wc_params* p);
int wc_default_step_params(wc_step_params* sp);
/* Fluid::setup constants without touching a device (Fluid.cpp:206-216). */
int wc_derive(const wc_params* p, wc_derived* d);
/* Number of CUDA devices this process sees (0 and WC_ERR_NO_DEVICE without a driver). */
int wc_device_count(int32_t* count);

/* Fluid::setup (Fluid.cpp:203-235) + Sort::prepareBuffers (Sort.cpp:67-94): allocate the
 * two particle buffers, count/offset/sorted buffers and scratch on p->device. */
typedef struct wc_handle wc_handle;
int wc_create(const wc_params* p, wc_handle** out);
/* Extended physics (see wc_physics): defaults with flags = 0; set / read back per handle.  Takes
 * effect from the next step on; WC_ERR_INVALID for unknown flag bits or non-finite values. */
int wc_default_physics(wc_physics* ph);
int wc_set_physics(wc_handle* h, const wc_physics* ph);
int wc_get_physics(const wc_handle* h, wc_physics* ph);
int wc_destroy(wc_handle* h);
int wc_get_derived(const wc_handle* h, wc_derived* d);

/* util::setParticles (util.cpp:59-63) into buffer 1, n <= capacity; sets num_particles. */
int wc_upload_particles(wc_handle* h, const wc_particle* host_aos, int32_t n);
/* util::getParticles (util.cpp:51-57): which = 1 (current state) or 2 (sorted, rho/P). */
int wc_download_particles(wc_handle* h, int32_t which, wc_particle* host_aos);

/* Fluid::update(double time) (Fluid.cpp:342-354): sort(buf1->buf2); density(buf2);
 * update(buf2->buf1, dt = frame_dt * time_scale).  Asynchronous on the handle's stream. */
int wc_step(wc_handle* h, float frame_dt, const wc_step_params* sp);
/* The same step with HOST particle buffers on both sides, for callers that keep the state on
 * the host (util::setParticles -> Fluid::update -> util::getParticles, util.cpp:42-63, as one
 * call): uploads n particles from host_in (NULL: step the resident state), steps, and leaves
 * the new buffer 1 in host_out (n * 32 bytes, cell-sorted order like wc_download_particles(1)).
 * When host_out is page-locked (cudaHostAlloc / cudaHostRegister) the update kernel stores
 * each result straight into it over PCIe, so the device-to-host copy overlaps the kernel
 * instead of following it; pageable memory takes the copy path.  Synchronous: host_out is
 * complete on return. */
int wc_step_host(wc_handle* h, float frame_dt, const wc_step_params* sp,
                 const wc_particle* host_in, int32_t n, wc_particle* host_out);
/* The same step with the new buffer 1 ALSO written, by the update kernel itself, as the
 * reference's 32-byte AoS Particle records (util.h:29-35) into aos_dst: a device pointer --
 * e.g. the SSBO Fluid::renderParticles binds at index 0 (Fluid.cpp:389-406,
 * assets/fluid/particle.vert:19-31), registered with cudaGraphicsGLRegisterBuffer and mapped
 * -- or page-locked mapped host memory.  num_particles * 32 bytes, cell-sorted order.  No pack
 * kernel, no extra pass over the state.  Asynchronous on the handle's stream. */
int wc_step_export(wc_handle* h, float frame_dt, const wc_step_params* sp, void* aos_dst);
/* Stage-level entry points, like the reference's separate runXProg methods. */
int wc_sort_only(wc_handle* h);                                      /* Sort::run, Sort.cpp:254 */
int wc_density_only(wc_handle* h, const wc_step_params* sp);         /* Fluid.cpp:268 */
int wc_update_only(wc_handle* h, float frame_dt, const wc_step_params* sp); /* Fluid.cpp:294 */

/* Fluid::runAdvectProg (Fluid.cpp:326-337) on buffer 1, in place: advect.comp:20-59.  Dead in
 * the reference (its call is commented out, Fluid.cpp:351); provided so every shader under
 * assets/fluid has a counterpart.  Not part of wc_step. */
int wc_advect_only(wc_handle* h, float frame_dt);

/* util::getUints (util.cpp:65-71) for the sort outputs; any pointer may be NULL.
 * cell_ids[n], counts[num_bins], offsets[num_bins], sorted_perm[n], neighbour_counts[n]. */
int wc_download_cells(wc_handle* h, uint32_t* cell_ids, uint32_t* counts, uint32_t* offsets,
                      uint32_t* sorted_perm, uint32_t* neighbour_counts);
/* Total force F of update.comp:195 per particle of the last update (3 floats each);
 * needs WC_FLAG_DEBUG_OUTPUTS. */
int wc_download_forces(wc_handle* h, float* forces_xyz);

/* Replace density/pressure (and everything else) of buffer 2 from host AoS: lets a test
 * drive wc_update_only with arbitrary inputs, like binding an SSBO by hand. */
int wc_upload_sorted(wc_handle* h, const wc_particle* host_aos, int32_t n);

int wc_device_ptrs(wc_handle* h, wc_device_view* view);
/* The particle count alone (what `numItems` is to Sort, Sort.h:24).  Unlike wc_device_ptrs this
 * hands out no pointers: wc_device_ptrs lets the caller write buffer 1 behind the library's
 * back, so after it every sort hashes the positions afresh instead of using the cell counts the
 * previous update pass prepared. */
int wc_get_num_particles(wc_handle* h, int32_t* n);
/* Pack buffer `which` (1 or 2) as 32-byte AoS into a DEVICE buffer (n * 32 bytes). */
int wc_export_aos_device(wc_handle* h, int32_t which, void* device_dst);
int wc_sync(wc_handle* h);

/* ---- z-slab decomposition (new; the reference is single-GPU) --------------
 * One step on a slab handle is the call sequence below; the caller moves the listed device
 * buffers between neighbouring ranks (NCCL / any transport) between the calls.  Direction
 * index 0 = the rank below (lower z), 1 = the rank above.  A missing neighbour's receive
 * buffers must be zero-filled (wc_slab_clear_recv).
 *   wc_slab_sort_count   unpack mig_in, hash + count + scan of the owned layers
 *     exchange lc_send[d] -> neighbour's lc_recv[1-d]              (lc_bytes each)
 *   wc_slab_sync_info    host sync; returns the particle counts of this step (a host-driven
 *                        transport needs them to size the halo messages; optional otherwise)
 *   wc_slab_reorder      ghost-layer offsets + stable reorder of the owned particles
 *     exchange halo positions: first / last owned layer of buffer 2 -> neighbour's ghost slots
 *   wc_slab_density
 *     exchange halo positions (now carrying density) and velocities/pressures
 *   wc_slab_update       force + integrate, then extraction of particles that left the slab
 *     exchange mig_out[d] -> neighbour's mig_in[1-d]               (mig_bytes each)
 * Concatenating the ranks' buffers in z order gives bit-identical results to one
 * whole-grid handle (tests/test_slab_gpu.py). */
typedef struct wc_slab_view {
    void* mig_out[2];
    void* mig_in[2];
    void* lc_send[2];
    void* lc_recv[2];
    void* pos_rho_sorted;  /* float4 buffer 2; owned particles start at index owned_first */
    void* vel_pres_sorted;
    uint64_t mig_bytes;    /* size of one migrant message: 32-byte header + capacity * 32 */
    uint64_t lc_bytes;     /* size of one layer-count message: 4 * (2 + grid_res^2) */
    int32_t owned_first;   /* = slab_ghost_capacity */
    int32_t reserved;
} wc_slab_view;

/* info[8] = { n_owned, n_first_layer, n_last_layer, n_ghost_below, n_ghost_above,
 *             errors (capacity overflows / lost migrants, sticky), migrants_in_below,
 *             migrants_in_above } */
int wc_slab_get_view(wc_handle* h, wc_slab_view* view);
int wc_slab_clear_recv(wc_handle* h, int32_t direction);
int wc_slab_sort_count(wc_handle* h);
int wc_slab_sync_info(wc_handle* h, int32_t info[8]);
int wc_slab_reorder(wc_handle* h);
int wc_slab_density(wc_handle* h, const wc_step_params* sp);
int wc_slab_update(wc_handle* h, float frame_dt, const wc_step_params* sp);

/* ---- peer-memory exchange: the slab step without a host-driven transport ----------------
 * After the two neighbours of a slab handle are attached, the wc_slab_* phase calls move the
 * four per-step messages themselves: each phase copies its output into the neighbour's
 * buffers (mapped peer memory, NVLink) on the handle's stream and raises a step counter in
 * the neighbour's signal array; the consuming phase first queues a one-thread kernel that
 * waits for that counter.  The caller then simply issues, per step,
 *   wc_slab_sort_count, wc_slab_sync_info, wc_slab_reorder, wc_slab_density, wc_slab_update
 * on every rank, with no exchange of its own.  Attach before the first step.
 *   same process (tests, one process driving several GPUs): wc_slab_peer_attach
 *   one process per GPU: wc_slab_ipc_export -> ship the 456-byte blob to the neighbour by any
 *   means (bench.py uses torch.distributed's object all-gather) -> wc_slab_peer_open.
 * A missing neighbour is simply not attached (and wc_slab_clear_recv zeroes its buffers). */
#define WC_IPC_HANDLE_BYTES 64
typedef struct wc_slab_ipc {
    /* cudaIpcMemHandle_t of: buffer 2 positions, buffer 2 velocities, mig_in[0], mig_in[1],
     * lc_recv[0], lc_recv[1], signal array */
    unsigned char mem[7][WC_IPC_HANDLE_BYTES];
    int32_t device;
    int32_t ghost_capacity;
} wc_slab_ipc;
int wc_slab_ipc_export(wc_handle* h, wc_slab_ipc* out);
/* One whole step of a slab handle whose neighbours are all attached (or absent): the five
 * phase calls above in one.  Nothing in the step depends on a host read-back -- the step's
 * counts live in a device record, every launch is sized by capacity, waits and signals sit
 * inside the consuming / producing kernels -- so with info == NULL the call only queues the
 * step (eleven kernels) and returns; one host thread can then drive every GPU of the box
 * (core::Fluid::devices, wc_headless --gpus N) and run steps ahead.  With info != NULL it
 * waits for the step and reports its counts (as wc_slab_sync_info); a non-zero error word
 * (capacity overflow, lost migrants, a neighbour's signal missing) returns WC_ERR_CAPACITY.
 * Errors are sticky on the device: an asynchronous caller meets them at its next info read,
 * download or wc_device_ptrs. */
int wc_slab_step_peer(wc_handle* h, float frame_dt, const wc_step_params* sp, int32_t info[8]);
/* wc_step_host for a slab handle (util::setParticles -> update -> util::getParticles as one
 * call per rank): uploads n particles from host_in (NULL: step the resident state), steps, and
 * leaves the slab's new buffer 1 in host_out (room for out_capacity >= `capacity` particles;
 * info[0] of them are valid).  Page-locked host_out: the update kernel stores the records into
 * it directly.  Synchronous; info must not be NULL. */
int wc_slab_step_peer_host(wc_handle* h, float frame_dt, const wc_step_params* sp,
                           const wc_particle* host_in, int32_t n, wc_particle* host_out,
                           int32_t out_capacity, int32_t info[8]);
int wc_slab_peer_open(wc_handle* h, int32_t direction, const wc_slab_ipc* peer);
int wc_slab_peer_attach(wc_handle* h, int32_t direction, wc_handle* peer);

/* ---- inspection ----------------------------------------------------------
 * The reference verifies itself by eye: render modes 1-4 colour a particle red when it left
 * the box or its density is not positive (assets/fluid/particle.vert:35-55), and
 * Sort::printGrids (Sort.cpp:237-249) / util::printParticles (util.cpp:113-126) dump tables
 * to the debug console.  wc_diagnose reduces buffer `which` (1 = current state, 2 = sorted
 * input of the last step) on the device -- one pass over 32 bytes per particle, fp64 sums
 * folded in a fixed order (same buffer -> same bits) -- and syncs.  Sums cover the VALID
 * particles only.  Works on slab handles too (the owned particles; cell fields are -1). */
#define WC_DIAG_HIST_BINS 32
#define WC_DIAG_HIST_PER_UNIT 4 /* histogram bins per unit of density / rest_density */
typedef struct wc_diagnostics {
    int64_t particles;       /* particles looked at */
    int64_t invalid;         /* outside [0,size]^3, non-finite state, or density <= 0 */
    int64_t out_of_box;      /* the position part of `invalid` alone */
    int64_t at_speed_clamp;  /* |v| component at MAX_SPEED = 50 (update.comp:5,199) */
    double mass;             /* valid particles x particleMass (Fluid.cpp:210) */
    double momentum[3];      /* m sum v */
    double kinetic_energy;   /* m/2 sum |v|^2 */
    double centre_of_mass[3];
    double max_speed;
    double density_min, density_max, density_mean;
    double pressure_min, pressure_max, pressure_mean;
    /* density / rest_density in bins of 1/WC_DIAG_HIST_PER_UNIT: [0,.25) [.25,.5) ... the
     * last bin also takes everything above */
    int64_t density_hist[WC_DIAG_HIST_BINS];
    int64_t max_cell_count;  /* largest bin of the last sort, -1 if there is none */
    int64_t nonempty_cells;  /* bins holding at least one particle, -1 likewise */
} wc_diagnostics;
int wc_diagnose(wc_handle* h, int32_t which, float rest_density, wc_diagnostics* out);

/* Milliseconds per stage of the last wc_step (needs WC_FLAG_STAGE_TIMING; syncs). */
int wc_stage_times(wc_handle* h, float ms[WC_NUM_STAGES]);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int wc_launch_count(const wc_handle* h, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif
