// wc_capi.cu -- the C-ABI of include/wc_sph.h: handle, buffers, stage orchestration.
//
// Host-side counterpart of the reference's Fluid::setup / Fluid::update
// (src/core/Fluid.cpp:203-235, :342-354) and Sort::prepareBuffers / Sort::run
// (src/core/Sort.cpp:67-94, :254-267).  No CPU fallback anywhere in this file.

#include "../../include/wc_sph.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "wc_common.cuh"
#include "wc_sort.cuh"
#include "wc_sph_tile.cuh"
#include "wc_sph_v1.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define WC_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver       \
                            ? WC_ERR_NO_DEVICE                                             \
                            : WC_ERR_CUDA,                                                 \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__,  \
                        __LINE__);                                                         \
    } while (0)

#define WC_CHECK_LAUNCH(h) \
    do {                   \
        (h)->launches++;   \
        WC_CUDA(cudaGetLastError()); \
    } while (0)

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Smallest fp32 x with sqrtf(x) >= h.  sqrtf is correctly rounded and monotone, so
// "sqrt(d2) >= h" (density.comp:117) is exactly "d2 >= T".
float dist2_threshold(float h) {
    float x = h * h;
    while (sqrtf(x) >= h && x > 0.0f) x = nextafterf(x, 0.0f);
    while (!(sqrtf(x) >= h)) x = nextafterf(x, INFINITY);
    return x;
}

// GLSL min/max semantics, as in the oracle.
inline float gmin(float x, float y) { return y < x ? y : x; }
inline float gmax(float x, float y) { return x < y ? y : x; }

// update.comp:105-113,118-121: all operands are uniforms, so it is evaluated once here.
bool mouse_ray_hits_box(const wc_step_params& sp, float size) {
    float t1[3], t2[3];
    for (int a = 0; a < 3; a++) {
        const float tmin = (0.0f - sp.mouse_origin[a]) / sp.mouse_dir[a];
        const float tmax = (size - sp.mouse_origin[a]) / sp.mouse_dir[a];
        t1[a] = gmin(tmin, tmax);
        t2[a] = gmax(tmin, tmax);
    }
    const float tnear = gmax(gmax(t1[0], t1[1]), t1[2]);
    const float tfar = gmin(gmin(t2[0], t2[1]), t2[2]);
    return !(tnear > tfar);
}

}  // namespace

struct wc_handle {
    wc_params p;
    wc_derived d;
    int n = 0;         // current particle count
    int cap = 0;
    int num_bins = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;

    float4* pos[2] = {nullptr, nullptr};  // [0] = buffer 1 (state), [1] = buffer 2 (sorted)
    float4* vel[2] = {nullptr, nullptr};
    float4* aos = nullptr;                // staging for the 32-byte AoS surface
    uint32_t* cell_ids = nullptr;
    uint32_t* ranks = nullptr;
    uint32_t* ids = nullptr;              // arrival-ordered IDs (sort.comp's raw output)
    uint32_t* perm = nullptr;             // stable permutation: perm[dst] = src
    uint32_t* offsets = nullptr;          // num_bins + 1
    uint32_t* neighbour_counts = nullptr;
    float4* forces = nullptr;
    // density -> update neighbour list (wc_sph_tile.cuh NbrList)
    uint32_t* nbr_idx = nullptr;
    uint32_t* nbr_mask = nullptr;
    uint32_t* nbr_words = nullptr;
    int nbr_cap_words = 0;
    bool nbr_valid = false;

    // Arena cleared once per sort: counts | scan status | scan tile counter.
    void* arena = nullptr;
    size_t arena_bytes = 0;
    uint32_t* counts = nullptr;
    unsigned long long* scan_status = nullptr;
    unsigned int* scan_counter = nullptr;

    cudaEvent_t ev[WC_NUM_STAGES + 1] = {};
    bool have_times = false;
    bool sorted_valid = false;
};

namespace {

using namespace wc;

SphConsts make_consts(const wc_handle* h, const wc_step_params& sp, float frame_dt) {
    SphConsts c;
    c.n = h->n;
    c.G = h->p.grid_res;
    c.bin = h->d.bin_size;
    c.size = h->p.size;
    c.h = h->d.kernel_radius;
    c.h2 = c.h * c.h;
    c.T = h->d.dist2_threshold;
    c.m = h->d.particle_mass;
    c.poly6C = h->d.poly6_const;
    c.spikyC = h->d.spiky_const;
    c.viscC = h->d.visc_const;
    c.mu = sp.viscosity_coefficient;
    c.k = sp.stiffness;
    c.rho0 = sp.rest_density;
    c.P0 = sp.rest_pressure;
    for (int a = 0; a < 3; a++) {
        c.g[a] = sp.gravity[a];
        c.mo[a] = sp.mouse_origin[a];
        c.md[a] = sp.mouse_dir[a];
    }
    c.dt = frame_dt * h->p.time_scale;  // Fluid.cpp:308
    c.mouse_hits = mouse_ray_hits_box(sp, h->p.size) ? 1 : 0;
    return c;
}

int record(wc_handle* h, int idx) {
    if (h->p.flags & WC_FLAG_STAGE_TIMING) WC_CUDA(cudaEventRecord(h->ev[idx], h->stream));
    return WC_OK;
}

// Sort::run (Sort.cpp:254-267).
int run_sort(wc_handle* h, bool timed) {
    const int n = h->n, G = h->p.grid_res;
    const float bin = h->d.bin_size;
    int rc;
    if (timed && (rc = record(h, 0))) return rc;
    // clearCountBuffer (Sort.cpp:255) -- one memset also resets the scan bookkeeping.
    WC_CUDA(cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
    if (n > 0) {
        k_hash_count<<<div_up(n, 256), 256, 0, h->stream>>>(h->pos[0], n, bin, G, h->cell_ids,
                                                             h->ranks, h->counts);
        WC_CHECK_LAUNCH(h);
    }
    if (timed && (rc = record(h, 1))) return rc;
    k_scan<<<div_up(h->num_bins, kScanTile), kScanThreads, 0, h->stream>>>(
        h->counts, h->offsets, h->num_bins, h->scan_status, h->scan_counter);
    WC_CHECK_LAUNCH(h);
    if (timed && (rc = record(h, 2))) return rc;
    if (n > 0) {
        k_scatter_ids<<<div_up(n, 256), 256, 0, h->stream>>>(h->cell_ids, h->ranks, h->offsets, n,
                                                              h->ids);
        WC_CHECK_LAUNCH(h);
        k_reorder<<<div_up(n, 256), 256, 0, h->stream>>>(h->ids, h->offsets, n, bin, G, h->pos[0],
                                                          h->vel[0], h->pos[1], h->vel[1],
                                                          h->perm);
        WC_CHECK_LAUNCH(h);
    }
    if (timed && (rc = record(h, 3))) return rc;
    h->sorted_valid = true;
    h->nbr_valid = false;
    return WC_OK;
}

int run_density(wc_handle* h, const wc_step_params& sp) {
    if (h->n == 0) return WC_OK;
    const SphConsts c = make_consts(h, sp, 0.0f);
    const bool dbg = h->p.flags & WC_FLAG_DEBUG_OUTPUTS;
    const NbrList list{h->nbr_idx, h->nbr_mask, h->nbr_words, h->nbr_cap_words};
    int rc = (h->p.flags & WC_FLAG_SIMPLE_KERNELS)
                 ? -1
                 : launch_density_tile(h->pos[1], h->vel[1], h->offsets, c,
                                       dbg ? h->neighbour_counts : nullptr, list, h->stream);
    h->nbr_valid = (rc == 0) && h->nbr_idx != nullptr;
    if (rc == -1) {  // geometry the tile kernel does not cover: simple path
        if (dbg)
            k_density_v1<true><<<div_up(h->n, 128), 128, 0, h->stream>>>(
                h->pos[1], h->vel[1], h->offsets, c, h->neighbour_counts);
        else
            k_density_v1<false><<<div_up(h->n, 128), 128, 0, h->stream>>>(
                h->pos[1], h->vel[1], h->offsets, c, nullptr);
    }
    WC_CHECK_LAUNCH(h);
    return WC_OK;
}

int run_update(wc_handle* h, const wc_step_params& sp, float frame_dt) {
    if (h->n == 0) return WC_OK;
    const SphConsts c = make_consts(h, sp, frame_dt);
    const bool dbg = h->p.flags & WC_FLAG_DEBUG_OUTPUTS;
    // The list is only trusted when the density pass that built it saw these positions.
    const NbrList list = h->nbr_valid
                             ? NbrList{h->nbr_idx, h->nbr_mask, h->nbr_words, h->nbr_cap_words}
                             : NbrList{nullptr, nullptr, nullptr, 0};
    int rc = (h->p.flags & WC_FLAG_SIMPLE_KERNELS)
                 ? -1
                 : launch_update_tile(h->pos[1], h->vel[1], h->offsets, c, h->pos[0], h->vel[0],
                                      dbg ? h->forces : nullptr, list, h->stream);
    if (rc == -1) {
        if (dbg)
            k_update_v1<true><<<div_up(h->n, 128), 128, 0, h->stream>>>(
                h->pos[1], h->vel[1], h->offsets, c, h->pos[0], h->vel[0], h->forces);
        else
            k_update_v1<false><<<div_up(h->n, 128), 128, 0, h->stream>>>(
                h->pos[1], h->vel[1], h->offsets, c, h->pos[0], h->vel[0], nullptr);
    }
    WC_CHECK_LAUNCH(h);
    return WC_OK;
}

int check_step_params(const wc_step_params* sp) {
    if (!sp) return fail(WC_ERR_INVALID, "step params are NULL");
    return WC_OK;
}

}  // namespace

extern "C" {

int wc_abi_version(void) { return WC_ABI_VERSION; }

const char* wc_last_error(void) { return g_err; }

int wc_default_params(wc_params* p) {
    if (!p) return fail(WC_ERR_INVALID, "params is NULL");
    std::memset(p, 0, sizeof(*p));
    p->num_particles = 80000;   // Fluid.cpp:12
    p->capacity = 0;
    p->grid_res = 21;           // Fluid.cpp:14
    p->size = 1.0f;             // Fluid.cpp:10
    p->particle_radius = 0.01f; // Fluid.cpp:17
    p->time_scale = 0.012f;     // Fluid.cpp:24
    p->device = 0;
    p->flags = 0;
    p->neighbour_list_words = 0;
    p->stream = nullptr;
    return WC_OK;
}

int wc_default_step_params(wc_step_params* sp) {
    if (!sp) return fail(WC_ERR_INVALID, "step params is NULL");
    std::memset(sp, 0, sizeof(*sp));
    sp->viscosity_coefficient = 200.0f;  // Fluid.cpp:19
    sp->stiffness = 100.0f;              // Fluid.cpp:20
    sp->rest_density = 500.0f;           // Fluid.cpp:18
    sp->rest_pressure = 0.0f;            // Fluid.cpp:21
    sp->gravity[0] = 0.0f;
    sp->gravity[1] = -1.0f * 900.0f;     // Fluid.cpp:15-16
    sp->gravity[2] = 0.0f;
    // mouse_ray_ is uninitialised in the reference before the first mouse move (Q19);
    // default to a ray that misses the box so the mouse force is exactly zero.
    sp->mouse_origin[0] = sp->mouse_origin[1] = sp->mouse_origin[2] = -10.0f;
    sp->mouse_dir[0] = -1.0f;
    return WC_OK;
}

int wc_derive(const wc_params* p, wc_derived* d) {
    if (!p || !d) return fail(WC_ERR_INVALID, "NULL argument");
    if (p->grid_res < 1 || p->grid_res > 1290)  // G^3 must fit in int32
        return fail(WC_ERR_INVALID, "grid_res %d out of range [1, 1290]", p->grid_res);
    if (!(p->size > 0.0f) || !(p->particle_radius > 0.0f))
        return fail(WC_ERR_INVALID, "size and particle_radius must be positive");
    // Fluid.cpp:207-216 (constants evaluated in double from the float radius, then cast)
    d->num_bins = p->grid_res * p->grid_res * p->grid_res;
    d->bin_size = p->size / (float)p->grid_res;
    d->kernel_radius = p->particle_radius * 4.0f;
    d->particle_mass = p->particle_radius * 8.0f;
    const double h = (double)d->kernel_radius, pi = 3.14159265358979323846;
    d->poly6_const = (float)(315.0 / (64.0 * pi * std::pow(h, 9)));
    d->spiky_const = (float)(-45.0 / (pi * std::pow(h, 6)));
    d->visc_const = (float)(45.0 / (pi * std::pow(h, 6)));
    d->dist2_threshold = dist2_threshold(d->kernel_radius);
    return WC_OK;
}

int wc_create(const wc_params* p, wc_handle** out) {
    if (!p || !out) return fail(WC_ERR_INVALID, "NULL argument");
    *out = nullptr;
    wc_derived d;
    int rc = wc_derive(p, &d);
    if (rc) return rc;
    if (p->num_particles < 0) return fail(WC_ERR_INVALID, "num_particles < 0");
    const int cap = p->capacity > 0 ? p->capacity : (p->num_particles > 0 ? p->num_particles : 1);
    if (cap < p->num_particles) return fail(WC_ERR_CAPACITY, "capacity < num_particles");
    if (d.bin_size < d.kernel_radius)
        return fail(WC_ERR_INVALID,
                    "binSize %g < kernelRadius %g: the 27-cell stencil would miss neighbours "
                    "(Fluid.cpp:13)", d.bin_size, d.kernel_radius);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(WC_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (p->device < 0 || p->device >= ndev)
        return fail(WC_ERR_NO_DEVICE, "device %d out of range (%d devices)", p->device, ndev);
    WC_CUDA(cudaSetDevice(p->device));
    cudaDeviceProp prop;
    WC_CUDA(cudaGetDeviceProperties(&prop, p->device));
    if (prop.major != 10)
        return fail(WC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    p->device, prop.major, prop.minor);

    wc_handle* h = new (std::nothrow) wc_handle();
    if (!h) return fail(WC_ERR_INVALID, "out of host memory");
    h->p = *p;
    h->d = d;
    h->n = p->num_particles;
    h->cap = cap;
    h->num_bins = d.num_bins;
    if (p->stream) {
        h->stream = (cudaStream_t)p->stream;
    } else {
        e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete h;
            return fail(WC_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        h->own_stream = true;
    }

    const size_t capz = (size_t)cap;
    const size_t nb = (size_t)d.num_bins;
    const size_t counts_bytes = ((nb + 1) * sizeof(uint32_t) + 255) / 256 * 256;
    const size_t tiles = (nb + kScanTile - 1) / kScanTile + 1;
    const size_t status_bytes = (tiles * sizeof(unsigned long long) + 255) / 256 * 256;
    h->arena_bytes = counts_bytes + status_bytes + 256;

#define WC_ALLOC(ptr, bytes)                                                            \
    do {                                                                                \
        e = cudaMalloc((void**)&(ptr), (bytes));                                        \
        if (e != cudaSuccess) {                                                         \
            wc_destroy(h);                                                              \
            return fail(WC_ERR_CUDA, "cudaMalloc(%zu bytes) for %s: %s", (size_t)(bytes), #ptr, \
                        cudaGetErrorString(e));                                         \
        }                                                                               \
    } while (0)

    for (int b = 0; b < 2; b++) {
        WC_ALLOC(h->pos[b], capz * sizeof(float4));
        WC_ALLOC(h->vel[b], capz * sizeof(float4));
    }
    WC_ALLOC(h->aos, capz * 2 * sizeof(float4));
    WC_ALLOC(h->cell_ids, capz * sizeof(uint32_t));
    WC_ALLOC(h->ranks, capz * sizeof(uint32_t));
    WC_ALLOC(h->ids, capz * sizeof(uint32_t));
    WC_ALLOC(h->perm, capz * sizeof(uint32_t));
    WC_ALLOC(h->offsets, (nb + 1) * sizeof(uint32_t));
    WC_ALLOC(h->arena, h->arena_bytes);
    if (p->neighbour_list_words >= 0 && !(p->flags & WC_FLAG_SIMPLE_KERNELS)) {
        h->nbr_cap_words = p->neighbour_list_words > 0 ? p->neighbour_list_words : 32;
        const size_t warps = (size_t)tile_warps(cap);
        WC_ALLOC(h->nbr_idx, warps * h->nbr_cap_words * 32 * sizeof(uint32_t));
        WC_ALLOC(h->nbr_mask, warps * h->nbr_cap_words * 32 * sizeof(uint32_t));
        WC_ALLOC(h->nbr_words, warps * sizeof(uint32_t));
    }
    if (p->flags & WC_FLAG_DEBUG_OUTPUTS) {
        WC_ALLOC(h->neighbour_counts, capz * sizeof(uint32_t));
        WC_ALLOC(h->forces, capz * sizeof(float4));
    }
#undef WC_ALLOC
    h->counts = (uint32_t*)h->arena;
    h->scan_status = (unsigned long long*)((char*)h->arena + counts_bytes);
    h->scan_counter = (unsigned int*)((char*)h->arena + counts_bytes + status_bytes);

    for (int b = 0; b < 2; b++) {
        cudaMemsetAsync(h->pos[b], 0, capz * sizeof(float4), h->stream);
        cudaMemsetAsync(h->vel[b], 0, capz * sizeof(float4), h->stream);
    }
    cudaMemsetAsync(h->offsets, 0, (nb + 1) * sizeof(uint32_t), h->stream);
    cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream);
    if (p->flags & WC_FLAG_STAGE_TIMING) {
        for (int i = 0; i <= WC_NUM_STAGES; i++) {
            e = cudaEventCreate(&h->ev[i]);
            if (e != cudaSuccess) {
                wc_destroy(h);
                return fail(WC_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
            }
        }
    }
    e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        wc_destroy(h);
        return fail(WC_ERR_CUDA, "initial clear failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return WC_OK;
}

int wc_destroy(wc_handle* h) {
    if (!h) return WC_OK;
    cudaSetDevice(h->p.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int b = 0; b < 2; b++) {
        cudaFree(h->pos[b]);
        cudaFree(h->vel[b]);
    }
    cudaFree(h->aos);
    cudaFree(h->cell_ids);
    cudaFree(h->ranks);
    cudaFree(h->ids);
    cudaFree(h->perm);
    cudaFree(h->offsets);
    cudaFree(h->arena);
    cudaFree(h->neighbour_counts);
    cudaFree(h->forces);
    cudaFree(h->nbr_idx);
    cudaFree(h->nbr_mask);
    cudaFree(h->nbr_words);
    for (int i = 0; i <= WC_NUM_STAGES; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return WC_OK;
}

int wc_get_derived(const wc_handle* h, wc_derived* d) {
    if (!h || !d) return fail(WC_ERR_INVALID, "NULL argument");
    *d = h->d;
    return WC_OK;
}

static int upload_into(wc_handle* h, int buf, const wc_particle* host_aos, int32_t n) {
    if (!h || (!host_aos && n > 0)) return fail(WC_ERR_INVALID, "NULL argument");
    if (n < 0) return fail(WC_ERR_INVALID, "n < 0");
    if (n > h->cap) return fail(WC_ERR_CAPACITY, "n = %d exceeds capacity %d", n, h->cap);
    WC_CUDA(cudaSetDevice(h->p.device));
    h->n = n;
    if (n > 0) {
        WC_CUDA(cudaMemcpyAsync(h->aos, host_aos, (size_t)n * sizeof(wc_particle),
                                cudaMemcpyHostToDevice, h->stream));
        wc::k_aos_to_soa<<<div_up(n, 256), 256, 0, h->stream>>>(h->aos, n, h->pos[buf], h->vel[buf]);
        WC_CHECK_LAUNCH(h);
    }
    return WC_OK;
}

int wc_upload_particles(wc_handle* h, const wc_particle* host_aos, int32_t n) {
    int rc = upload_into(h, 0, host_aos, n);
    if (rc == WC_OK) h->sorted_valid = false;
    return rc;
}

int wc_upload_sorted(wc_handle* h, const wc_particle* host_aos, int32_t n) {
    if (h && n != h->n) return fail(WC_ERR_INVALID, "n = %d differs from num_particles %d", n, h->n);
    if (h) h->nbr_valid = false;  // positions may have changed under the neighbour list
    return upload_into(h, 1, host_aos, n);
}

int wc_export_aos_device(wc_handle* h, int32_t which, void* device_dst) {
    if (!h || !device_dst) return fail(WC_ERR_INVALID, "NULL argument");
    if (which != 1 && which != 2) return fail(WC_ERR_INVALID, "which must be 1 or 2");
    WC_CUDA(cudaSetDevice(h->p.device));
    if (h->n > 0) {
        wc::k_soa_to_aos<<<div_up(h->n, 256), 256, 0, h->stream>>>(
            h->pos[which - 1], h->vel[which - 1], h->n, (float4*)device_dst);
        WC_CHECK_LAUNCH(h);
    }
    return WC_OK;
}

int wc_download_particles(wc_handle* h, int32_t which, wc_particle* host_aos) {
    if (!h || (!host_aos && h->n > 0)) return fail(WC_ERR_INVALID, "NULL argument");
    int rc = wc_export_aos_device(h, which, h->aos);
    if (rc) return rc;
    if (h->n > 0)
        WC_CUDA(cudaMemcpyAsync(host_aos, h->aos, (size_t)h->n * sizeof(wc_particle),
                                cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return WC_OK;
}

int wc_step(wc_handle* h, float frame_dt, const wc_step_params* sp) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    int rc = check_step_params(sp);
    if (rc) return rc;
    WC_CUDA(cudaSetDevice(h->p.device));
    if ((rc = run_sort(h, true))) return rc;          // Fluid.cpp:347
    if ((rc = run_density(h, *sp))) return rc;        // Fluid.cpp:349
    if ((rc = record(h, 4))) return rc;
    if ((rc = run_update(h, *sp, frame_dt))) return rc;  // Fluid.cpp:350
    if ((rc = record(h, 5))) return rc;
    h->have_times = (h->p.flags & WC_FLAG_STAGE_TIMING) != 0;
    return WC_OK;
}

int wc_sort_only(wc_handle* h) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    WC_CUDA(cudaSetDevice(h->p.device));
    h->have_times = false;
    return run_sort(h, false);
}

int wc_density_only(wc_handle* h, const wc_step_params* sp) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    int rc = check_step_params(sp);
    if (rc) return rc;
    if (!h->sorted_valid) return fail(WC_ERR_INVALID, "wc_density_only needs a preceding sort");
    WC_CUDA(cudaSetDevice(h->p.device));
    h->have_times = false;
    return run_density(h, *sp);
}

int wc_update_only(wc_handle* h, float frame_dt, const wc_step_params* sp) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    int rc = check_step_params(sp);
    if (rc) return rc;
    if (!h->sorted_valid) return fail(WC_ERR_INVALID, "wc_update_only needs a preceding sort");
    WC_CUDA(cudaSetDevice(h->p.device));
    h->have_times = false;
    return run_update(h, *sp, frame_dt);
}

int wc_download_cells(wc_handle* h, uint32_t* cell_ids, uint32_t* counts, uint32_t* offsets,
                      uint32_t* sorted_perm, uint32_t* neighbour_counts) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    WC_CUDA(cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->n, nb = (size_t)h->num_bins;
    if (neighbour_counts && !h->neighbour_counts)
        return fail(WC_ERR_INVALID, "neighbour counts need WC_FLAG_DEBUG_OUTPUTS");
    if (cell_ids && n)
        WC_CUDA(cudaMemcpyAsync(cell_ids, h->cell_ids, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (counts)
        WC_CUDA(cudaMemcpyAsync(counts, h->counts, nb * 4, cudaMemcpyDeviceToHost, h->stream));
    if (offsets)
        WC_CUDA(cudaMemcpyAsync(offsets, h->offsets, nb * 4, cudaMemcpyDeviceToHost, h->stream));
    if (sorted_perm && n)
        WC_CUDA(cudaMemcpyAsync(sorted_perm, h->perm, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (neighbour_counts && n)
        WC_CUDA(cudaMemcpyAsync(neighbour_counts, h->neighbour_counts, n * 4,
                                cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return WC_OK;
}

int wc_download_forces(wc_handle* h, float* forces_xyz) {
    if (!h || !forces_xyz) return fail(WC_ERR_INVALID, "NULL argument");
    if (!h->forces) return fail(WC_ERR_INVALID, "forces need WC_FLAG_DEBUG_OUTPUTS");
    WC_CUDA(cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->n;
    if (n == 0) return WC_OK;
    float4* tmp = (float4*)malloc(n * sizeof(float4));
    if (!tmp) return fail(WC_ERR_INVALID, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(tmp, h->forces, n * sizeof(float4), cudaMemcpyDeviceToHost,
                                    h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        free(tmp);
        return fail(WC_ERR_CUDA, "force download: %s", cudaGetErrorString(e));
    }
    for (size_t i = 0; i < n; i++) {
        forces_xyz[3 * i + 0] = tmp[i].x;
        forces_xyz[3 * i + 1] = tmp[i].y;
        forces_xyz[3 * i + 2] = tmp[i].z;
    }
    free(tmp);
    return WC_OK;
}

int wc_device_ptrs(wc_handle* h, wc_device_view* v) {
    if (!h || !v) return fail(WC_ERR_INVALID, "NULL argument");
    for (int b = 0; b < 2; b++) {
        v->pos_rho[b] = h->pos[b];
        v->vel_pres[b] = h->vel[b];
    }
    v->cell_ids = h->cell_ids;
    v->counts = h->counts;
    v->offsets = h->offsets;
    v->sorted = h->perm;
    v->neighbour_counts = h->neighbour_counts;
    v->forces = h->forces;
    v->stream = (void*)h->stream;
    v->num_particles = h->n;
    v->capacity = h->cap;
    return WC_OK;
}

int wc_sync(wc_handle* h) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    WC_CUDA(cudaSetDevice(h->p.device));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return WC_OK;
}

int wc_stage_times(wc_handle* h, float ms[WC_NUM_STAGES]) {
    if (!h || !ms) return fail(WC_ERR_INVALID, "NULL argument");
    if (!(h->p.flags & WC_FLAG_STAGE_TIMING) || !h->have_times)
        return fail(WC_ERR_INVALID, "no timed wc_step yet (WC_FLAG_STAGE_TIMING)");
    WC_CUDA(cudaSetDevice(h->p.device));
    WC_CUDA(cudaEventSynchronize(h->ev[WC_NUM_STAGES]));
    for (int s = 0; s < WC_NUM_STAGES; s++)
        WC_CUDA(cudaEventElapsedTime(&ms[s], h->ev[s], h->ev[s + 1]));
    return WC_OK;
}

int wc_launch_count(const wc_handle* h, uint64_t* launches) {
    if (!h || !launches) return fail(WC_ERR_INVALID, "NULL argument");
    *launches = h->launches;
    return WC_OK;
}

}  // extern "C"
