// wc_capi.cu -- the C-ABI of include/wc_sph.h: handle, buffers, stage orchestration.
//
// Host-side counterpart of the reference's Fluid::setup / Fluid::update
// (src/core/Fluid.cpp:203-235, :342-354) and Sort::prepareBuffers / Sort::run
// (src/core/Sort.cpp:67-94, :254-267).  No CPU fallback anywhere in this file.

#include "../../include/wc_sph.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "wc_common.cuh"
#include "wc_diag.cuh"
#include "wc_slab.cuh"
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
    uint32_t* big_cells = nullptr;        // cells above kBigCell particles (k_reorder_big)
    uint32_t* big_count = nullptr;        // (in the arena: zero at the start of every sort)
    int big_cap = 0;
    uint32_t* offsets = nullptr;          // num_bins + 1
    uint32_t* neighbour_counts = nullptr;
    float4* forces = nullptr;
    wc_physics phys = {};                 // extended physics (flags 0 = the reference's step)
    // density -> update neighbour list (wc_sph_tile.cuh NbrList)
    uint32_t* nbr_idx = nullptr;
    uint32_t* nbr_mask = nullptr;
    uint32_t* nbr_words = nullptr;
    int nbr_cap_words = 0;
    bool nbr_valid = false;
    // group table of the tiled gather kernels (wc_sph_tile.cuh k_build_groups)
    uint4* groups = nullptr;
    uint32_t* num_groups = nullptr;
    int groups_cap = 0;

    // Arena cleared once per sort: counts | scan status | scan tile counter.
    void* arena = nullptr;
    size_t arena_bytes = 0;
    size_t counts_bytes = 0, status_bytes = 0;  // layout of the arena's head (bind_arena)
    size_t x_off0 = 0, xbytes[4] = {};          // ... and of the slab mode's extra scan states
    // Pre-hash (tiled kernels; whole grid and z-slabs): the update pass counts the NEXT sort's cells into a
    // second arena -- and writes its cell ids / arrival ranks into second arrays -- on the
    // positions it has just integrated; if buffer 1 is still what the update stored when the
    // next sort starts, that sort swaps the sets in and begins at the scan.
    void* arena_next = nullptr;
    uint32_t* cell_ids_next = nullptr;
    uint32_t* ranks_next = nullptr;
    bool prehash_valid = false;   // arena_next / *_next describe buffer 1 as it is now
    bool prehash_off = false;     // raw device pointers were handed out (wc_device_ptrs)
    uint32_t* counts = nullptr;
    unsigned long long* scan_status = nullptr;
    unsigned int* scan_counter = nullptr;

    wc::DiagPartial* diag = nullptr;      // wc_diagnose scratch: [0] result, [1..] block partials
    unsigned int* diag_cells = nullptr;   // [2]

    cudaEvent_t ev[WC_NUM_STAGES + 1] = {};
    bool have_times = false;
    bool sorted_valid = false;

    // ---- z-slab mode (wc_slab.cuh); all zero / false for a whole-grid handle
    bool slab = false;
    int z_begin = 0, z_end = 0;  // owned global z-layers [z_begin, z_end)
    int Lz = 0;                  // layers of the local table: owned + 2 ghost layers (or G)
    int zbase = 0;               // global layer of local layer 0
    int M = 0;                   // migrant slots on either side of buffer 1's owned region
    int Cg = 0;                  // ghost slots on either side of buffer 2's owned region
    int n_first = 0, n_last = 0, n_glow = 0, n_ghigh = 0;  // of the last step the host looked at
    int m_in_host[2] = {0, 0};
    // The step's own counts live on the device (wc::SlabDyn); the host copies above and `n` are
    // refreshed only when somebody asks (slab_sync_host), so a step never waits for the host.
    wc::SlabDyn* dyn = nullptr;
    bool host_stale = false;     // steps were queued since the host copies were refreshed
    // Asynchronous steps run at most kRunAhead steps ahead of the device: the launch queue of a
    // stream is finite, and a host thread that blocks in a launch on one handle can no longer
    // queue the work a neighbouring handle of the same thread is waiting for.
    static constexpr int kRunAhead = 4;
    cudaEvent_t step_done[kRunAhead] = {};
    uint32_t ghost_step = 0;     // step_no whose ghost tables have been queued
    uint32_t* done = nullptr;    // [8] last-block counters of the signalling kernels (in arena)
    float4* mig_out[2] = {nullptr, nullptr};   // [0] to rank-1, [1] to rank+1 (header + AoS)
    float4* mig_in[2] = {nullptr, nullptr};    // [0] from rank-1, [1] from rank+1
    uint32_t* lc_send[2] = {nullptr, nullptr}; // layer-count messages [n, G*G counts]
    uint32_t* lc_recv[2] = {nullptr, nullptr};
    uint32_t* info_host = nullptr;             // pinned: this step's counts (k_ghost_tables)
    unsigned long long* scan_status_x[4] = {}; // extra scan states: ghost-low, ghost-high, mig 0/1
    unsigned int* scan_counter_x[4] = {};
    size_t mig_bytes = 0, lc_bytes = 0;

    // ---- peer-memory exchange (wc_slab_peer_*): the neighbours' buffers mapped into this
    // process / context, their signal arrays, and this handle's own signal array
    struct Peer {
        bool on = false, ipc = false;
        float4* pos1 = nullptr;     // neighbour's buffer 2 (sorted), incl. its ghost slots
        float4* vel1 = nullptr;
        float4* mig_in = nullptr;   // neighbour's mig_in[1 - d]
        uint32_t* lc_recv = nullptr;  // neighbour's lc_recv[1 - d]
        uint32_t* sig = nullptr;    // neighbour's signal array
        void* ipc_base[5] = {};     // what cudaIpcCloseMemHandle needs
        int n_owned = 0;            // neighbour's owned count of the current step
    } peer[2];
    uint32_t* sig = nullptr;        // [kSigPhases * 2]: raised by the neighbours
    uint32_t step_no = 0;
    bool peer_mode() const { return peer[0].on || peer[1].on; }
    // A neighbour on the SAME device (virtual ranks of the tests): a kernel whose every block
    // spins for that neighbour's signal could keep the neighbour's own kernels off the SMs, so
    // the waits then stay stand-alone one-block kernels in front of the consumers.
    bool same_device_peer = false;
    bool fused_waits() const { return peer_mode() && !same_device_peer; }
};

enum { kSigLc = 0, kSigHaloPos = 1, kSigHaloRho = 2, kSigMig = 3, kSigPhases = 4 };

namespace {

using namespace wc;

// The pointers into the head of the (current) arena.
void bind_arena(wc_handle* h) {
    char* a = (char*)h->arena;
    h->counts = (uint32_t*)a;
    h->scan_status = (unsigned long long*)(a + h->counts_bytes);
    h->scan_counter = (unsigned int*)(a + h->counts_bytes + h->status_bytes);
    h->num_groups = (uint32_t*)(a + h->counts_bytes + h->status_bytes + 128);
    h->big_count = (uint32_t*)(a + h->counts_bytes + h->status_bytes + 192);
    if (h->slab) {
        size_t off = h->x_off0;
        for (int k = 0; k < 4; k++) {
            h->scan_status_x[k] = (unsigned long long*)(a + off);
            h->scan_counter_x[k] = (unsigned int*)(a + off + h->xbytes[k] - 256);
            off += h->xbytes[k];
        }
        h->done = (uint32_t*)(a + off);  // 8 block counters, zero at every sort
    }
}

SphConstsExt make_consts(const wc_handle* h, const wc_step_params& sp, float frame_dt) {
    SphConstsExt c;
    c.n = h->n;
    c.first = h->Cg;
    c.G = h->p.grid_res;
    c.Gz = h->Lz;
    c.zbase = h->zbase;
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
    // extended physics (oracle make_consts derives the same floats)
    const wc_physics& ph = h->phys;
    c.phys = ph.flags;
    c.sigma = ph.surface_tension;
    c.n_min = ph.surface_threshold;
    c.grad_m = (-6.0f * c.poly6C) * c.m;
    c.h2x3 = 3.0f * c.h2;
    c.wall_acc = ph.wall_stiffness / (c.dt * c.dt);
    c.wall_d = ph.wall_distance;
    const float wallC = (float)(0.78539816339744830962 * (double)c.poly6C * std::pow((double)c.h, 9.0));
    c.wall_w = (ph.wall_rest_density > 0.0f ? ph.wall_rest_density : sp.rest_density) * wallC;
    return c;
}

int record(wc_handle* h, int idx) {
    if (h->p.flags & WC_FLAG_STAGE_TIMING) WC_CUDA(cudaEventRecord(h->ev[idx], h->stream));
    return WC_OK;
}

enum { kDoneInfo = 0, kDoneGhost = 1, kDoneSort = 2, kDoneDensity = 3, kDoneMigrants = 4,
       kDoneStandAlone = -2 };

// What a slab kernel gets beyond its arrays (wc_common.cuh SlabRef): the device record, the
// attached neighbours' buffer 2, and -- peer mode only -- the local flags it waits on
// (wait_phase, -1: none) and the neighbours' flags its last block raises (raise_phase, with
// its block counter `done_slot`).  `value` is the step number waited for / raised.
SlabRef slab_ref(const wc_handle* h, int wait_phase = -1, int raise_phase = -1, int done_slot = -1,
                 long long value = -1) {
    SlabRef r = SlabRef();
    if (!h->slab) return r;
    r.dyn = h->dyn;
    r.Cg = (uint32_t)h->Cg;
    r.cap = (uint32_t)h->cap;
    r.step_no = value >= 0 ? (uint32_t)value : h->step_no;
    // (the block counter only matters when somebody is signalled -- or to close the step)
    if (done_slot >= 0 && (h->peer_mode() || raise_phase < 0 || done_slot == kDoneMigrants))
        r.done = h->done + done_slot;
    for (int d = 0; d < 2; d++) {
        if (!h->peer[d].on) continue;
        r.peer_pos[d] = h->peer[d].pos1;
        r.peer_vel[d] = h->peer[d].vel1;
        if (wait_phase >= 0 && (h->fused_waits() || done_slot == kDoneStandAlone))
            r.wait[d] = h->sig + wait_phase * 2 + d;
        // the neighbour in direction d sees this rank as its direction 1 - d
        if (raise_phase >= 0) r.raise[d] = h->peer[d].sig + raise_phase * 2 + (1 - d);
    }
    return r;
}

// The host copies of the step's counts, on demand: waits for the stream, then reads the
// page-locked record k_ghost_tables stored and the sticky error word.
int slab_sync_host(wc_handle* h, bool report = true) {
    if (!h->slab || !h->host_stale) return WC_OK;
    uint32_t err = 0;
    WC_CUDA(cudaMemcpyAsync(&err, &h->dyn->errors, sizeof(err), cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    const uint32_t* hi = h->info_host;
    h->n = (int)hi[0];
    h->n_first = (int)hi[1], h->n_last = (int)hi[2];
    h->n_glow = (int)hi[3], h->n_ghigh = (int)hi[4];
    h->m_in_host[0] = (int)hi[6], h->m_in_host[1] = (int)hi[7];
    h->peer[0].n_owned = (int)hi[8], h->peer[1].n_owned = (int)hi[9];
    h->info_host[5] = err;
    h->host_stale = false;
    if (h->n > h->cap) h->n = h->cap;  // the step is dead (kSlabErrOwned); keep host copies in range
    // an asynchronous step cannot report: its sticky error word is met here, at the next look
    if (err && report)
        return fail(err & kSlabErrTimeout ? WC_ERR_CUDA : WC_ERR_CAPACITY,
                    "a queued slab step failed (error bits 0x%x:%s%s%s%s%s%s); counts of the failing "
                    "step: %u owned (capacity %d), boundary layers %u / %u, ghosts %u / %u (ghost "
                    "capacity %d), migrant capacity %d", err,
                    err & kSlabErrOwned ? " owned-capacity" : "", err & kSlabErrGhost ? " ghost-capacity" : "",
                    err & kSlabErrHalo ? " halo-layer>neighbour-ghost-capacity" : "",
                    err & kSlabErrMigrants ? " migrant-capacity" : "",
                    err & kSlabErrStray ? " stray-migrant(moved>1-layer)" : "",
                    err & kSlabErrTimeout ? " neighbour-signal-timeout" : "", hi[0], h->cap, hi[1],
                    hi[2], hi[3], hi[4], h->Cg, h->M);
    return WC_OK;
}

// Host view of where this rank's halo layers live in the attached neighbours: only for the
// simple cross-check kernels (WC_FLAG_SIMPLE_KERNELS), whose density pass has no remote
// stores, so the halo goes as explicit peer copies.  Needs slab_sync_host.
int peer_copy_halo(wc_handle* h) {
    for (int d = 0; d < 2; d++) {
        if (!h->peer[d].on) continue;
        const size_t n_layer = (size_t)(d == 0 ? h->n_first : h->n_last);
        const size_t src = (size_t)h->Cg + (d == 0 ? 0 : (size_t)(h->n - h->n_last));
        const size_t dst = d == 0 ? (size_t)h->Cg + h->peer[0].n_owned : (size_t)h->Cg - h->n_last;
        if (n_layer == 0) continue;
        WC_CUDA(cudaMemcpyAsync(h->peer[d].pos1 + dst, h->pos[1] + src, n_layer * sizeof(float4),
                                cudaMemcpyDeviceToDevice, h->stream));
        WC_CUDA(cudaMemcpyAsync(h->peer[d].vel1 + dst, h->vel[1] + src, n_layer * sizeof(float4),
                                cudaMemcpyDeviceToDevice, h->stream));
    }
    return WC_OK;
}

// Stand-alone wait and / or signal, a one-block kernel: for the simple cross-check kernels,
// which carry neither, and for the waits of handles whose neighbour shares the device.
int slab_sync_kernel(wc_handle* h, int wait_phase, int raise_phase, int done_slot, long long value = -1) {
    if (!h->slab || !h->peer_mode()) return WC_OK;
    SlabRef r = slab_ref(h, wait_phase, raise_phase, done_slot >= 0 ? done_slot : kDoneStandAlone, value);
    k_slab_sync<<<1, 32, 0, h->stream>>>(r);
    WC_CHECK_LAUNCH(h);
    return WC_OK;
}

// The wait in front of a consumer kernel when it cannot ride inside the kernel.
int slab_pre_wait(wc_handle* h, int phase, long long value = -1) {
    if (!h->slab || !h->peer_mode() || h->fused_waits()) return WC_OK;
    return slab_sync_kernel(h, phase, -1, -1, value);
}

// Sort::run part 1 (Sort.cpp:255-259): clear, count, scan.  In slab mode the input is the
// virtual array [migrants from below | owned | migrants from above] and only the owned
// layers are scanned (the ghost layers' offsets come from the neighbours' counts).
int sort_count_phase(wc_handle* h, bool timed) {
    const int G = h->p.grid_res;
    const float bin = h->d.bin_size;
    int rc;
    if (timed && (rc = record(h, 0))) return rc;
    // The update pass of the previous step may already have hashed and counted its output
    // (PreHash, wc_sph_tile.cuh): then its arena and arrays are swapped in and the step starts at
    // the scan (slab mode: after hashing only the received migrants).
    const bool prehashed = h->prehash_valid;
    h->prehash_valid = false;
    if (prehashed) {
        std::swap(h->arena, h->arena_next);
        std::swap(h->cell_ids, h->cell_ids_next);
        std::swap(h->ranks, h->ranks_next);
        bind_arena(h);
    } else {
        // clearCountBuffer (Sort.cpp:255) -- one memset also resets the scan bookkeeping.
        WC_CUDA(cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
    }
    if (!h->slab) {
        if (h->n > 0 && !prehashed) {
            k_hash_count<<<div_up(h->n, 256), 256, 0, h->stream>>>(h->pos[0], h->n, bin, G,
                                                                    h->cell_ids, h->ranks,
                                                                    h->counts);
            WC_CHECK_LAUNCH(h);
        }
        if (timed && (rc = record(h, 1))) return rc;
        k_scan<<<div_up(h->num_bins, kScanTile), kScanThreads, 0, h->stream>>>(
            h->counts, h->offsets, h->num_bins, h->scan_status, h->scan_counter, 0u);
        WC_CHECK_LAUNCH(h);
        if (timed && (rc = record(h, 2))) return rc;
        return WC_OK;
    }
    const int G2 = G * G, M = h->M;
    h->step_no++;
    h->host_stale = true;
    // (waits for, and) unpacks last step's migrants -> the slots before / after the owned region
    if ((rc = slab_pre_wait(h, kSigMig, (long long)h->step_no - 1))) return rc;
    k_slab_begin<<<div_up(2 * M, 256), 256, 0, h->stream>>>(
        h->mig_in[0], h->mig_in[1], M, h->pos[0], h->vel[0],
        slab_ref(h, kSigMig, -1, -1, (long long)h->step_no - 1));
    WC_CHECK_LAUNCH(h);
    // (pre-hashed: the owned slots are done, only the two migrant windows remain)
    k_hash_count_slab<<<div_up(prehashed ? 2 * M : M + h->cap + M, 256), 256, 0, h->stream>>>(
        h->pos[0], M, h->dyn, bin, G, h->z_begin, h->z_end, h->cell_ids, h->ranks, h->counts,
        prehashed ? 1 : 0);
    WC_CHECK_LAUNCH(h);
    if (timed && (rc = record(h, 1))) return rc;
    const int owned_bins = (h->Lz - 2) * G2;
    k_scan<<<div_up(owned_bins, kScanTile), kScanThreads, 0, h->stream>>>(
        h->counts + G2, h->offsets + G2, owned_bins, h->scan_status, h->scan_counter,
        (uint32_t)h->Cg);
    WC_CHECK_LAUNCH(h);
    // counts into the slab record; layer counts to the neighbours (stored there directly in
    // peer mode, whose last block raises the flags)
    k_slab_info<<<div_up(G2, 256), 256, 0, h->stream>>>(
        h->counts, h->offsets, G2, h->Lz, h->lc_send[0], h->lc_send[1],
        h->peer[0].on ? h->peer[0].lc_recv : nullptr, h->peer[1].on ? h->peer[1].lc_recv : nullptr,
        slab_ref(h, -1, kSigLc, kDoneInfo));
    WC_CHECK_LAUNCH(h);
    if (timed && (rc = record(h, 2))) return rc;
    return WC_OK;
}

// The ghost layers of the table from the neighbours' layer-count messages (queued once per step).
int slab_ghost_phase(wc_handle* h) {
    if (h->ghost_step == h->step_no) return WC_OK;
    const int G2 = h->p.grid_res * h->p.grid_res;
    int rcw = slab_pre_wait(h, kSigLc);
    if (rcw) return rcw;
    k_ghost_tables<<<2, kScanThreads, 0, h->stream>>>(h->lc_recv[0], h->lc_recv[1], G2, h->Lz,
                                                     h->counts, h->offsets, h->info_host,
                                                     slab_ref(h, kSigLc, -1, kDoneGhost));
    WC_CHECK_LAUNCH(h);
    h->ghost_step = h->step_no;
    return WC_OK;
}

// Sort::run part 2 (Sort.cpp:263-264): the stable reorder into buffer 2.  n_in / n_sorted:
// sizes of the input and of the sorted array -- in slab mode launch bounds only (the kernels
// read the true counts from the slab record).
int sort_reorder_phase(wc_handle* h, bool timed, int n_in, int n_sorted) {
    const int G = h->p.grid_res;
    const float bin = h->d.bin_size;
    int rc;
    const bool moved = n_in > 0 && n_sorted > 0;
    const ReorderIO io{h->pos[0], h->vel[0], h->pos[1] + h->Cg, h->vel[1] + h->Cg, h->perm};
    if (moved) {
        k_scatter_ids<<<div_up(n_in, 256), 256, 0, h->stream>>>(
            h->cell_ids, h->ranks, h->offsets, n_in, h->ids, (uint32_t)h->Cg, slab_ref(h), h->M);
        WC_CHECK_LAUNCH(h);
    }
    // slab mode: the ghost layers of the table -- which wait for the neighbours' layer counts --
    // go here, behind the ID scatter (it only needs the owned layers), so that wait is covered
    if (h->slab && (rc = slab_ghost_phase(h))) return rc;
    if (moved) {
        k_reorder<<<div_up(n_sorted, 256), 256, 0, h->stream>>>(
            h->ids, h->offsets, n_sorted, bin, G, h->zbase, (uint32_t)h->Cg, io, slab_ref(h),
            h->big_cells, h->big_count, (uint32_t)h->big_cap);
        WC_CHECK_LAUNCH(h);
    }
    {   // group table for the gathers + the cells above kBigCell particles, one launch; in peer
        // mode its last block tells the neighbours that the halo positions are in their ghost slots
        const int row0 = h->slab ? G : 0, row1 = h->slab ? (h->Lz - 1) * G : h->Lz * G;
        int bits = 1;  // IDs index the (virtual) input: that many radix passes
        while (bits < 32 && (1ll << bits) < (long long)n_in) bits++;
        if (h->groups || moved || h->slab) {
            k_finish_sort<<<finish_sort_blocks(h->groups ? row1 - row0 : 0), kBigThreads, 0,
                            h->stream>>>(h->offsets, G, row0, row1, h->groups, h->num_groups,
                                         moved ? h->ids : nullptr, h->ranks, (uint32_t)h->Cg, io,
                                         slab_ref(h, -1, kSigHaloPos, kDoneSort), h->big_cells,
                                         h->big_count, (uint32_t)h->big_cap, (bits + 7) / 8);
            WC_CHECK_LAUNCH(h);
        }
    }
    if (timed && (rc = record(h, 3))) return rc;
    h->sorted_valid = true;
    h->nbr_valid = false;
    return WC_OK;
}

GroupTable group_table(const wc_handle* h) {
    // launch bound: from the particle count the host knows -- in slab mode the capacity
    return GroupTable{h->groups, h->num_groups,
                      max_groups(h->slab ? h->cap : h->n, (long long)h->Lz * h->p.grid_res)};
}

int run_sort(wc_handle* h, bool timed) {
    int rc = sort_count_phase(h, timed);
    if (rc) return rc;
    return sort_reorder_phase(h, timed, h->n, h->n);
}

int run_density(wc_handle* h, const wc_step_params& sp) {
    if (!h->slab && h->n == 0) return WC_OK;
    const SphConstsExt c = make_consts(h, sp, 0.0f);
    const bool dbg = h->p.flags & WC_FLAG_DEBUG_OUTPUTS;
    const NbrList list{h->nbr_idx, h->nbr_mask, h->nbr_words, h->nbr_cap_words};
    const bool simple = h->p.flags & WC_FLAG_SIMPLE_KERNELS;
    if (!simple)
        launch_density_tile(h->pos[1], h->vel[1], h->offsets, c, group_table(h),
                            dbg ? h->neighbour_counts : nullptr, list, h->stream,
                            slab_ref(h, kSigHaloPos, kSigHaloRho, kDoneDensity));
    h->nbr_valid = !simple && h->nbr_idx != nullptr;
    if (simple) {
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

int run_update(wc_handle* h, const wc_step_params& sp, float frame_dt,
               float4* aos_out = nullptr, bool prehash = true) {
    h->prehash_valid = false;  // buffer 1 is about to change
    if (!h->slab && h->n == 0) return WC_OK;
    const SphConstsExt c = make_consts(h, sp, frame_dt);
    const bool dbg = h->p.flags & WC_FLAG_DEBUG_OUTPUTS;
    // The list is only trusted when the density pass that built it saw these positions.
    const NbrList list = h->nbr_valid
                             ? NbrList{h->nbr_idx, h->nbr_mask, h->nbr_words, h->nbr_cap_words}
                             : NbrList{nullptr, nullptr, nullptr, 0};
    const bool simple = h->p.flags & WC_FLAG_SIMPLE_KERNELS;
    if (!simple) {
        PreHash pre{nullptr, nullptr, nullptr};
        if (prehash && h->arena_next && !h->prehash_off) {
            WC_CUDA(cudaMemsetAsync(h->arena_next, 0, h->arena_bytes, h->stream));
            // (indices of buffer 1's virtual array: the owned region starts at slot M)
            pre = PreHash{(uint32_t*)h->arena_next, h->cell_ids_next + h->M, h->ranks_next + h->M};
        }
        launch_update_tile(h->pos[1], h->vel[1], h->offsets, c, group_table(h), h->pos[0] + h->M,
                           h->vel[0] + h->M, dbg ? h->forces : nullptr, list, h->stream, aos_out,
                           slab_ref(h, kSigHaloRho), pre);
        h->prehash_valid = pre.counts != nullptr;
    }
    if (simple) {
        if (dbg)
            k_update_v1<true><<<div_up(h->n, 128), 128, 0, h->stream>>>(
                h->pos[1], h->vel[1], h->offsets, c, h->pos[0] + h->M, h->vel[0] + h->M,
                h->forces);
        else
            k_update_v1<false><<<div_up(h->n, 128), 128, 0, h->stream>>>(
                h->pos[1], h->vel[1], h->offsets, c, h->pos[0] + h->M, h->vel[0] + h->M, nullptr);
    }
    WC_CHECK_LAUNCH(h);
    return WC_OK;
}

int check_step_params(const wc_step_params* sp) {
    if (!sp) return fail(WC_ERR_INVALID, "step params are NULL");
    return WC_OK;
}

// wc_slab_update: force + integrate (optionally also stored as AoS records into aos_out, a
// device-addressable buffer), then the extraction of the particles that left the slab.
int slab_update_impl(wc_handle* h, float frame_dt, const wc_step_params* sp, float4* aos_out) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    if (!h->slab) return fail(WC_ERR_INVALID, "not a slab handle");
    WC_CUDA(cudaSetDevice(h->p.device));
    int rc = check_step_params(sp);
    if (rc) return rc;
    if (!h->sorted_valid) return fail(WC_ERR_INVALID, "wc_slab_update needs wc_slab_reorder");
    const float bin = h->d.bin_size;
    if (!(50.0f * fabsf(frame_dt * h->p.time_scale) < bin))
        return fail(WC_ERR_INVALID, "dt too large for one-layer migration: 50 * dt >= binSize");
    const bool simple = h->p.flags & WC_FLAG_SIMPLE_KERNELS;
    if (simple) {
        if ((rc = slab_sync_host(h))) return rc;
        if ((rc = slab_sync_kernel(h, kSigHaloRho, -1, -1))) return rc;
    } else if ((rc = slab_pre_wait(h, kSigHaloRho))) {
        return rc;
    }
    if ((rc = run_update(h, *sp, frame_dt, aos_out))) return rc;
    // Particles whose new z-layer left the slab: only the first / last owned layer can lose
    // any (|v| dt < binSize), and those layers are the head / tail of the sorted order.
    const int tiles = div_up(h->Cg, kScanTile) > 0 ? div_up(h->Cg, kScanTile) : 1;
    k_migrants<<<dim3(tiles, 2), kScanThreads, 0, h->stream>>>(
        h->pos[0] + h->M, h->vel[0] + h->M, bin, h->p.grid_res, h->z_begin, h->z_end, h->M,
        h->mig_out[0], h->mig_out[1], h->peer[0].on ? h->peer[0].mig_in : nullptr,
        h->peer[1].on ? h->peer[1].mig_in : nullptr, h->scan_status_x[2], h->scan_status_x[3],
        h->scan_counter_x[2], h->scan_counter_x[3], slab_ref(h, -1, kSigMig, kDoneMigrants));
    WC_CHECK_LAUNCH(h);
    h->host_stale = true;  // (the next step's input count changed on the device)
    if ((rc = record(h, 5))) return rc;
    h->have_times = (h->p.flags & WC_FLAG_STAGE_TIMING) != 0;
    return WC_OK;
}

// CUDA loads a kernel lazily, at its first launch, and that load can wait for running kernels.
// A slab step must never meet it: with the neighbour's work queued by the same host thread, a
// kernel spinning for that neighbour's signal would block the load that the rest of the queueing
// is waiting behind.  So a slab handle touches every kernel of the step once, at creation.
int preload_slab_kernels() {
    cudaFuncAttributes a;
    WC_CUDA(cudaFuncGetAttributes(&a, k_slab_begin));
    WC_CUDA(cudaFuncGetAttributes(&a, k_hash_count_slab));
    WC_CUDA(cudaFuncGetAttributes(&a, k_scan));
    WC_CUDA(cudaFuncGetAttributes(&a, k_slab_info));
    WC_CUDA(cudaFuncGetAttributes(&a, k_ghost_tables));
    WC_CUDA(cudaFuncGetAttributes(&a, k_scatter_ids));
    WC_CUDA(cudaFuncGetAttributes(&a, k_reorder));
    WC_CUDA(cudaFuncGetAttributes(&a, k_finish_sort));
    WC_CUDA(cudaFuncGetAttributes(&a, k_density_tile<false, true, false>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_density_tile<true, true, false>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_update_tile<false, true, false>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_update_tile<true, true, false>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_density_tile<false, true, true>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_density_tile<true, true, true>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_update_tile<false, true, true>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_update_tile<true, true, true>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_density_v1<false>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_density_v1<true>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_update_v1<false>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_update_v1<true>));
    WC_CUDA(cudaFuncGetAttributes(&a, k_migrants));
    WC_CUDA(cudaFuncGetAttributes(&a, k_slab_sync));
    WC_CUDA(cudaFuncGetAttributes(&a, k_aos_to_soa));
    WC_CUDA(cudaFuncGetAttributes(&a, k_soa_to_aos));
    return WC_OK;
}

// One whole step of a slab handle whose neighbours are attached (or absent): the five phases
// queued back to back.  info == NULL: returns without waiting for the GPU (nothing in the step
// depends on a host read-back).  info != NULL: waits and reports the step's counts.
int slab_step_impl(wc_handle* h, float frame_dt, const wc_step_params* sp, int32_t info[8],
                   float4* aos_out) {
    int rc;
    if (h && h->slab && !info) {  // bound the run-ahead: wait for the step kRunAhead steps back
        cudaEvent_t ev = h->step_done[(h->step_no + 1) % wc_handle::kRunAhead];
        if (ev) WC_CUDA(cudaEventSynchronize(ev));
    }
    if ((rc = wc_slab_sort_count(h))) return rc;
    if (h->p.flags & WC_FLAG_SIMPLE_KERNELS) {  // host-sized cross-check kernels: needs the counts
        int32_t tmp[8];
        if ((rc = wc_slab_sync_info(h, tmp))) return rc;
    }
    if ((rc = wc_slab_reorder(h))) return rc;
    if ((rc = wc_slab_density(h, sp))) return rc;
    if ((rc = slab_update_impl(h, frame_dt, sp, aos_out))) return rc;
    if (!info) {
        cudaEvent_t& ev = h->step_done[h->step_no % wc_handle::kRunAhead];
        if (!ev) WC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        WC_CUDA(cudaEventRecord(ev, h->stream));
        return WC_OK;
    }
    h->host_stale = true;
    if ((rc = slab_sync_host(h, false))) return rc;
    for (int k = 0; k < 8; k++) info[k] = (int32_t)h->info_host[k];
    if (info[5] != 0)
        return fail(WC_ERR_CAPACITY, "slab capacity overflow, lost migrants or a missing neighbour "
                                     "signal (error bits 0x%x)", (unsigned)info[5]);
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

int wc_default_physics(wc_physics* ph) {
    if (!ph) return fail(WC_ERR_INVALID, "physics is NULL");
    ph->flags = 0;
    ph->surface_tension = 50.0f;
    ph->surface_threshold = 7.0f;
    ph->wall_stiffness = 0.5f;
    ph->wall_distance = 0.01f;
    ph->wall_rest_density = 0.0f;
    return WC_OK;
}

int wc_set_physics(wc_handle* h, const wc_physics* ph) {
    if (!h || !ph) return fail(WC_ERR_INVALID, "NULL argument");
    if (ph->flags & ~(WC_PHYS_WALL_PARTICLES | WC_PHYS_SURFACE_TENSION))
        return fail(WC_ERR_INVALID, "unknown physics flag bits 0x%x", ph->flags);
    const float v[5] = {ph->surface_tension, ph->surface_threshold, ph->wall_stiffness,
                        ph->wall_distance, ph->wall_rest_density};
    for (float x : v)
        if (!std::isfinite(x)) return fail(WC_ERR_INVALID, "physics values must be finite");
    if (ph->surface_threshold < 0.0f || ph->wall_stiffness < 0.0f || ph->wall_distance < 0.0f)
        return fail(WC_ERR_INVALID, "surface_threshold, wall_stiffness and wall_distance must be >= 0");
    h->phys = *ph;
    return WC_OK;
}

int wc_get_physics(const wc_handle* h, wc_physics* ph) {
    if (!h || !ph) return fail(WC_ERR_INVALID, "NULL argument");
    *ph = h->phys;
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

int wc_device_count(int32_t* count) {
    if (!count) return fail(WC_ERR_INVALID, "NULL argument");
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    *count = e == cudaSuccess ? n : 0;
    if (e != cudaSuccess || n == 0)
        return fail(WC_ERR_NO_DEVICE, "no CUDA device: %s",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
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

    const bool slab = p->slab_ghost_capacity > 0;
    if (slab) {
        if (p->slab_z_begin < 0 || p->slab_z_end > p->grid_res || p->slab_z_end <= p->slab_z_begin)
            return fail(WC_ERR_INVALID, "slab layers [%d, %d) are not inside [0, %d)",
                        p->slab_z_begin, p->slab_z_end, p->grid_res);
        if (p->slab_migrant_capacity <= 0)
            return fail(WC_ERR_INVALID, "slab_migrant_capacity must be positive in slab mode");
    }

    wc_handle* h = new (std::nothrow) wc_handle();
    if (!h) return fail(WC_ERR_INVALID, "out of host memory");
    h->p = *p;
    h->d = d;
    wc_default_physics(&h->phys);
    h->n = p->num_particles;
    h->cap = cap;
    h->slab = slab;
    h->Lz = p->grid_res;
    if (slab) {
        h->z_begin = p->slab_z_begin;
        h->z_end = p->slab_z_end;
        h->Lz = h->z_end - h->z_begin + 2;
        h->zbase = h->z_begin - 1;
        h->M = p->slab_migrant_capacity;
        h->Cg = p->slab_ghost_capacity;
    }
    // bins of the (slab-local) cell table
    h->num_bins = h->Lz * p->grid_res * p->grid_res;
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
    const size_t nb = (size_t)h->num_bins;
    const size_t G2 = (size_t)p->grid_res * p->grid_res;
    const size_t in_slots = capz + 2 * (size_t)h->M;    // buffer 1: [M | owned | M]
    // buffer 2: [Cg | owned | Cg] + the slack the cull's unconditional loads may read into
    const size_t sorted_slots = capz + 2 * (size_t)h->Cg + (size_t)kCullOverread;
    auto scan_state_bytes = [](size_t elems) {
        const size_t tiles = (elems + kScanTile - 1) / kScanTile + 1;
        return (tiles * sizeof(unsigned long long) + 255) / 256 * 256 + 256;
    };
    const size_t counts_bytes = ((nb + 1) * sizeof(uint32_t) + 255) / 256 * 256;
    const size_t status_bytes = scan_state_bytes(nb) - 256;
    // extra scan states (slab): ghost-low, ghost-high tables; the two migrant compactions
    const size_t xbytes[4] = {scan_state_bytes(G2), scan_state_bytes(G2), scan_state_bytes(capz),
                              scan_state_bytes(capz)};
    h->arena_bytes = counts_bytes + status_bytes + 256;  // last 256: tile counter, num_groups
    const size_t x_off0 = h->arena_bytes;
    if (slab) h->arena_bytes += xbytes[0] + xbytes[1] + xbytes[2] + xbytes[3] + 256;

#define WC_ALLOC(ptr, bytes)                                                            \
    do {                                                                                \
        e = cudaMalloc((void**)&(ptr), (bytes));                                        \
        if (e != cudaSuccess) {                                                         \
            wc_destroy(h);                                                              \
            return fail(WC_ERR_CUDA, "cudaMalloc(%zu bytes) for %s: %s", (size_t)(bytes), #ptr, \
                        cudaGetErrorString(e));                                         \
        }                                                                               \
    } while (0)

    WC_ALLOC(h->pos[0], in_slots * sizeof(float4));
    WC_ALLOC(h->vel[0], in_slots * sizeof(float4));
    WC_ALLOC(h->pos[1], sorted_slots * sizeof(float4));
    WC_ALLOC(h->vel[1], sorted_slots * sizeof(float4));
    WC_ALLOC(h->aos, capz * 2 * sizeof(float4));
    WC_ALLOC(h->cell_ids, in_slots * sizeof(uint32_t));
    WC_ALLOC(h->ranks, in_slots * sizeof(uint32_t));
    WC_ALLOC(h->ids, capz * sizeof(uint32_t));
    WC_ALLOC(h->perm, capz * sizeof(uint32_t));
    h->big_cap = cap / kBigCell + 1;      // more cells than this cannot exceed kBigCell each
    WC_ALLOC(h->big_cells, (size_t)h->big_cap * sizeof(uint32_t));
    WC_ALLOC(h->offsets, (nb + 1) * sizeof(uint32_t));
    WC_ALLOC(h->arena, h->arena_bytes);
    if (!(p->flags & WC_FLAG_SIMPLE_KERNELS)) {
        h->groups_cap = max_groups(cap, (long long)h->Lz * p->grid_res);
        const size_t groups = (size_t)h->groups_cap;
        WC_ALLOC(h->groups, (groups + 64) * sizeof(uint4));  // + launch-bound round-up
        if (p->neighbour_list_words >= 0) {
            // Default capacity: ~3x the candidates a 32-particle group keeps at the scene's mean
            // number density (box of its targets grown by h), between 32 and 96 words.  A group
            // that overflows falls back to a fresh search in the update pass (still exact).
            const double box = (double)p->size, hh = (double)d.kernel_radius, bs = (double)d.bin_size;
            const double kept = (double)cap / (box * box * box) * (1.6 * bs + 2 * hh) *
                                (bs + 2 * hh) * (bs + 2 * hh);
            int words = (int)std::ceil(3.0 * kept / 32.0);
            words = words < 32 ? 32 : (words > 96 ? 96 : words);
            h->nbr_cap_words = p->neighbour_list_words > 0 ? p->neighbour_list_words : words;
            WC_ALLOC(h->nbr_idx, groups * h->nbr_cap_words * 32 * sizeof(uint32_t));
            WC_ALLOC(h->nbr_mask, groups * h->nbr_cap_words * 32 * sizeof(uint32_t));
            WC_ALLOC(h->nbr_words, groups * sizeof(uint32_t));
        }
    }
    if (p->flags & WC_FLAG_DEBUG_OUTPUTS) {
        WC_ALLOC(h->neighbour_counts, capz * sizeof(uint32_t));
        WC_ALLOC(h->forces, capz * sizeof(float4));
    }
    h->counts_bytes = counts_bytes;
    h->status_bytes = status_bytes;
    h->x_off0 = x_off0;
    for (int k = 0; k < 4; k++) h->xbytes[k] = xbytes[k];
    bind_arena(h);
    if (!(p->flags & WC_FLAG_SIMPLE_KERNELS)) {  // the pre-hash set (see wc_handle)
        WC_ALLOC(h->arena_next, h->arena_bytes);
        WC_ALLOC(h->cell_ids_next, in_slots * sizeof(uint32_t));
        WC_ALLOC(h->ranks_next, in_slots * sizeof(uint32_t));
    }
    if (slab) {
        h->mig_bytes = (size_t)(kMigHeaderFloat4 + 2 * (size_t)h->M) * sizeof(float4);
        h->lc_bytes = (kLcHeader + G2) * sizeof(uint32_t);
        for (int k = 0; k < 2; k++) {
            WC_ALLOC(h->mig_out[k], h->mig_bytes);
            WC_ALLOC(h->mig_in[k], h->mig_bytes);
            WC_ALLOC(h->lc_send[k], h->lc_bytes);
            WC_ALLOC(h->lc_recv[k], h->lc_bytes);
            cudaMemsetAsync(h->mig_out[k], 0, h->mig_bytes, h->stream);
            cudaMemsetAsync(h->mig_in[k], 0, h->mig_bytes, h->stream);
            cudaMemsetAsync(h->lc_send[k], 0, h->lc_bytes, h->stream);
            cudaMemsetAsync(h->lc_recv[k], 0, h->lc_bytes, h->stream);
        }
        WC_ALLOC(h->dyn, 256);
        cudaMemsetAsync(h->dyn, 0, 256, h->stream);
        WC_ALLOC(h->sig, 256);
        cudaMemsetAsync(h->sig, 0, 256, h->stream);
        e = cudaMallocHost((void**)&h->info_host, 64);
        if (e != cudaSuccess) {
            wc_destroy(h);
            return fail(WC_ERR_CUDA, "cudaMallocHost: %s", cudaGetErrorString(e));
        }
        std::memset(h->info_host, 0, 64);
        if ((rc = preload_slab_kernels())) {
            wc_destroy(h);
            return rc;
        }
    }

#undef WC_ALLOC
    cudaMemsetAsync(h->pos[0], 0, in_slots * sizeof(float4), h->stream);
    cudaMemsetAsync(h->vel[0], 0, in_slots * sizeof(float4), h->stream);
    cudaMemsetAsync(h->pos[1], 0, sorted_slots * sizeof(float4), h->stream);
    cudaMemsetAsync(h->vel[1], 0, sorted_slots * sizeof(float4), h->stream);
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
    cudaFree(h->big_cells);
    cudaFree(h->offsets);
    cudaFree(h->arena);
    cudaFree(h->arena_next);
    cudaFree(h->cell_ids_next);
    cudaFree(h->ranks_next);
    cudaFree(h->neighbour_counts);
    cudaFree(h->forces);
    cudaFree(h->nbr_idx);
    cudaFree(h->nbr_mask);
    cudaFree(h->nbr_words);
    cudaFree(h->groups);
    cudaFree(h->diag);
    cudaFree(h->diag_cells);
    for (int k = 0; k < 2; k++) {
        cudaFree(h->mig_out[k]);
        cudaFree(h->mig_in[k]);
        cudaFree(h->lc_send[k]);
        cudaFree(h->lc_recv[k]);
    }
    cudaFree(h->dyn);
    cudaFree(h->sig);
    for (int d = 0; d < 2; d++)
        if (h->peer[d].ipc)
            for (void* base : h->peer[d].ipc_base)
                if (base) cudaIpcCloseMemHandle(base);
    if (h->info_host) cudaFreeHost(h->info_host);
    for (cudaEvent_t ev : h->step_done)
        if (ev) cudaEventDestroy(ev);
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
    if (buf == 0) h->prehash_valid = false;
    if (h->slab) {  // the device record follows: buffer 1's owned count is the next step's input
        const int rcs = slab_sync_host(h);
        if (rcs) return rcs;
        h->n = n;
        const uint32_t un = (uint32_t)n;
        if (buf == 0)
            WC_CUDA(cudaMemcpyAsync(&h->dyn->n_in_old, &un, 4, cudaMemcpyHostToDevice, h->stream));
        else
            WC_CUDA(cudaMemcpyAsync(&h->dyn->n, &un, 4, cudaMemcpyHostToDevice, h->stream));
        WC_CUDA(cudaStreamSynchronize(h->stream));  // `un` is a stack variable
    }
    if (n > 0) {
        WC_CUDA(cudaMemcpyAsync(h->aos, host_aos, (size_t)n * sizeof(wc_particle),
                                cudaMemcpyHostToDevice, h->stream));
        const int first = buf == 0 ? h->M : h->Cg;  // owned region of buffer 1 / buffer 2
        wc::k_aos_to_soa<<<div_up(n, 256), 256, 0, h->stream>>>(h->aos, n, h->pos[buf] + first,
                                                              h->vel[buf] + first);
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
    if (h && h->slab) { int rcs = slab_sync_host(h); if (rcs) return rcs; }
    if (h && n != h->n) return fail(WC_ERR_INVALID, "n = %d differs from num_particles %d", n, h->n);
    if (h) h->nbr_valid = false;  // positions may have changed under the neighbour list
    return upload_into(h, 1, host_aos, n);
}

int wc_export_aos_device(wc_handle* h, int32_t which, void* device_dst) {
    if (!h || !device_dst) return fail(WC_ERR_INVALID, "NULL argument");
    if (which != 1 && which != 2) return fail(WC_ERR_INVALID, "which must be 1 or 2");
    WC_CUDA(cudaSetDevice(h->p.device));
    { int rcs = slab_sync_host(h); if (rcs) return rcs; }
    if (h->n > 0) {
        const int first = which == 1 ? h->M : h->Cg;
        wc::k_soa_to_aos<<<div_up(h->n, 256), 256, 0, h->stream>>>(
            h->pos[which - 1] + first, h->vel[which - 1] + first, h->n, (float4*)device_dst);
        WC_CHECK_LAUNCH(h);
    }
    return WC_OK;
}

int wc_download_particles(wc_handle* h, int32_t which, wc_particle* host_aos) {
    if (!h) return fail(WC_ERR_INVALID, "NULL argument");
    int rc = wc_export_aos_device(h, which, h->aos);  // (refreshes a slab handle's count first)
    if (rc == WC_OK && !host_aos && h->n > 0) return fail(WC_ERR_INVALID, "NULL argument");
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
    if (h->slab) return fail(WC_ERR_INVALID, "slab handle: use the wc_slab_* sequence");
    WC_CUDA(cudaSetDevice(h->p.device));
    if ((rc = run_sort(h, true))) return rc;          // Fluid.cpp:347
    if ((rc = run_density(h, *sp))) return rc;        // Fluid.cpp:349
    if ((rc = record(h, 4))) return rc;
    if ((rc = run_update(h, *sp, frame_dt))) return rc;  // Fluid.cpp:350
    if ((rc = record(h, 5))) return rc;
    h->have_times = (h->p.flags & WC_FLAG_STAGE_TIMING) != 0;
    return WC_OK;
}

int wc_step_export(wc_handle* h, float frame_dt, const wc_step_params* sp, void* aos_dst) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    int rc = check_step_params(sp);
    if (rc) return rc;
    if (h->slab) return fail(WC_ERR_INVALID, "slab handle: use the wc_slab_* sequence");
    if (!aos_dst && h->n > 0) return fail(WC_ERR_INVALID, "aos_dst is NULL");
    WC_CUDA(cudaSetDevice(h->p.device));
    // must be addressable from the device: device memory, or mapped page-locked host memory
    float4* dst = nullptr;
    if (h->n > 0) {
        cudaPointerAttributes attr{};
        const cudaError_t e = cudaPointerGetAttributes(&attr, aos_dst);
        cudaGetLastError();
        if (e != cudaSuccess || !attr.devicePointer ||
            (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeHost &&
             attr.type != cudaMemoryTypeManaged))
            return fail(WC_ERR_INVALID, "aos_dst is not device-addressable memory");
        dst = static_cast<float4*>(attr.devicePointer);
    }
    if ((rc = run_sort(h, true))) return rc;
    if ((rc = run_density(h, *sp))) return rc;
    if ((rc = record(h, 4))) return rc;
    if (h->p.flags & WC_FLAG_SIMPLE_KERNELS) {  // the cross-check kernels have no fused store
        if ((rc = run_update(h, *sp, frame_dt))) return rc;
        if ((rc = wc_export_aos_device(h, 1, dst))) return rc;
    } else if ((rc = run_update(h, *sp, frame_dt, dst))) {
        return rc;
    }
    if ((rc = record(h, 5))) return rc;
    h->have_times = (h->p.flags & WC_FLAG_STAGE_TIMING) != 0;
    return WC_OK;
}

int wc_step_host(wc_handle* h, float frame_dt, const wc_step_params* sp,
                 const wc_particle* host_in, int32_t n, wc_particle* host_out) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    int rc = check_step_params(sp);
    if (rc) return rc;
    if (h->slab) return fail(WC_ERR_INVALID, "slab handle: use the wc_slab_* sequence");
    if (host_in && (rc = wc_upload_particles(h, host_in, n))) return rc;
    if (!host_out && h->n > 0) return fail(WC_ERR_INVALID, "host_out is NULL");
    WC_CUDA(cudaSetDevice(h->p.device));
    // Page-locked destination: the update kernel writes the AoS records into it directly.
    float4* mapped = nullptr;
    if (h->n > 0 && !(h->p.flags & WC_FLAG_SIMPLE_KERNELS)) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, host_out) == cudaSuccess &&
            attr.type == cudaMemoryTypeHost && attr.devicePointer)
            mapped = static_cast<float4*>(attr.devicePointer);
        cudaGetLastError();  // an unregistered pointer is not an error here
    }
    if ((rc = run_sort(h, true))) return rc;
    if ((rc = run_density(h, *sp))) return rc;
    if ((rc = record(h, 4))) return rc;
    // (no pre-hash: the next call uploads a fresh buffer 1)
    if ((rc = run_update(h, *sp, frame_dt, mapped, false))) return rc;
    if ((rc = record(h, 5))) return rc;
    h->have_times = (h->p.flags & WC_FLAG_STAGE_TIMING) != 0;
    if (!mapped) return wc_download_particles(h, 1, host_out);
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return WC_OK;
}

int wc_sort_only(wc_handle* h) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    if (h->slab) return fail(WC_ERR_INVALID, "slab handle: use the wc_slab_* sequence");
    WC_CUDA(cudaSetDevice(h->p.device));
    // with WC_FLAG_STAGE_TIMING the three sort stages are timed (density / update read as 0)
    int rc = run_sort(h, true);
    if (rc) return rc;
    if ((rc = record(h, 4)) || (rc = record(h, 5))) return rc;
    h->have_times = (h->p.flags & WC_FLAG_STAGE_TIMING) != 0;
    return WC_OK;
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

int wc_advect_only(wc_handle* h, float frame_dt) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    if (h->slab) return fail(WC_ERR_INVALID, "wc_advect_only is not available on slab handles");
    WC_CUDA(cudaSetDevice(h->p.device));
    h->have_times = false;
    h->prehash_valid = false;
    if (h->n > 0) {
        k_advect<<<div_up(h->n, 256), 256, 0, h->stream>>>(h->pos[0], h->vel[0], h->n, h->p.size,
                                                           frame_dt * h->p.time_scale);
        WC_CHECK_LAUNCH(h);
    }
    return WC_OK;
}

int wc_download_cells(wc_handle* h, uint32_t* cell_ids, uint32_t* counts, uint32_t* offsets,
                      uint32_t* sorted_perm, uint32_t* neighbour_counts) {
    if (!h) return fail(WC_ERR_INVALID, "handle is NULL");
    WC_CUDA(cudaSetDevice(h->p.device));
    { int rcs = slab_sync_host(h); if (rcs) return rcs; }
    const size_t n = (size_t)h->n;
    const size_t G2 = (size_t)h->p.grid_res * h->p.grid_res;
    // slab mode: the owned layers only (skip the ghost-low layer), offsets relative to the
    // first owned slot; cell ids refer to the virtual input and are not exported.
    const size_t skip = h->slab ? G2 : 0, nb = (size_t)h->num_bins - 2 * skip;
    if (neighbour_counts && !h->neighbour_counts)
        return fail(WC_ERR_INVALID, "neighbour counts need WC_FLAG_DEBUG_OUTPUTS");
    if (cell_ids && h->slab) return fail(WC_ERR_INVALID, "cell_ids are not exported in slab mode");
    if (cell_ids && n)
        WC_CUDA(cudaMemcpyAsync(cell_ids, h->cell_ids, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (counts)
        WC_CUDA(cudaMemcpyAsync(counts, h->counts + skip, nb * 4, cudaMemcpyDeviceToHost,
                                h->stream));
    if (offsets)
        WC_CUDA(cudaMemcpyAsync(offsets, h->offsets + skip, nb * 4, cudaMemcpyDeviceToHost,
                                h->stream));
    if (sorted_perm && n)
        WC_CUDA(cudaMemcpyAsync(sorted_perm, h->perm, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (neighbour_counts && n)
        WC_CUDA(cudaMemcpyAsync(neighbour_counts, h->neighbour_counts, n * 4,
                                cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    if (offsets && h->slab)
        for (size_t b = 0; b < nb; b++) offsets[b] -= (uint32_t)h->Cg;
    return WC_OK;
}

int wc_download_forces(wc_handle* h, float* forces_xyz) {
    if (!h || !forces_xyz) return fail(WC_ERR_INVALID, "NULL argument");
    if (!h->forces) return fail(WC_ERR_INVALID, "forces need WC_FLAG_DEBUG_OUTPUTS");
    WC_CUDA(cudaSetDevice(h->p.device));
    { int rcs = slab_sync_host(h); if (rcs) return rcs; }
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

int wc_get_num_particles(wc_handle* h, int32_t* n) {
    if (!h || !n) return fail(WC_ERR_INVALID, "NULL argument");
    if (h->slab) {  // num_particles of a slab handle is device state: refresh the host copy
        WC_CUDA(cudaSetDevice(h->p.device));
        int rcs = slab_sync_host(h);
        if (rcs) return rcs;
    }
    *n = h->n;
    return WC_OK;
}

int wc_device_ptrs(wc_handle* h, wc_device_view* v) {
    if (!h || !v) return fail(WC_ERR_INVALID, "NULL argument");
    if (h->slab) {  // num_particles of a slab handle is device state: refresh the host copy
        WC_CUDA(cudaSetDevice(h->p.device));
        int rcs = slab_sync_host(h);
        if (rcs) return rcs;
    }
    // The caller may now write buffer 1 behind the library's back: from here on every sort
    // hashes the positions it finds (no pre-hash by the update pass).
    h->prehash_off = true;
    h->prehash_valid = false;
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

// ---------------------------------------------------------------------------- z-slab mode
#define WC_NEED_SLAB(h)                                                         \
    do {                                                                        \
        if (!(h)) return fail(WC_ERR_INVALID, "handle is NULL");                \
        if (!(h)->slab) return fail(WC_ERR_INVALID, "not a slab handle");       \
        WC_CUDA(cudaSetDevice((h)->p.device));                                  \
    } while (0)

int wc_slab_get_view(wc_handle* h, wc_slab_view* v) {
    if (!v) return fail(WC_ERR_INVALID, "NULL argument");
    WC_NEED_SLAB(h);
    for (int k = 0; k < 2; k++) {
        v->mig_out[k] = h->mig_out[k];
        v->mig_in[k] = h->mig_in[k];
        v->lc_send[k] = h->lc_send[k];
        v->lc_recv[k] = h->lc_recv[k];
    }
    v->pos_rho_sorted = h->pos[1];
    v->vel_pres_sorted = h->vel[1];
    v->mig_bytes = h->mig_bytes;
    v->lc_bytes = h->lc_bytes;
    v->owned_first = h->Cg;
    v->reserved = 0;
    return WC_OK;
}

int wc_slab_clear_recv(wc_handle* h, int32_t direction) {
    WC_NEED_SLAB(h);
    if (direction != 0 && direction != 1) return fail(WC_ERR_INVALID, "direction must be 0 or 1");
    WC_CUDA(cudaMemsetAsync(h->mig_in[direction], 0, 32, h->stream));
    WC_CUDA(cudaMemsetAsync(h->lc_recv[direction], 0, h->lc_bytes, h->stream));
    return WC_OK;
}

int wc_slab_sort_count(wc_handle* h) {
    WC_NEED_SLAB(h);
    h->have_times = false;
    return sort_count_phase(h, true);
}

int wc_slab_sync_info(wc_handle* h, int32_t info[8]) {
    if (!info) return fail(WC_ERR_INVALID, "NULL argument");
    WC_NEED_SLAB(h);
    int rc;
    if (h->step_no == 0) return fail(WC_ERR_INVALID, "wc_slab_sync_info before the first wc_slab_sort_count");
    if ((rc = slab_ghost_phase(h))) return rc;  // completes the record (needs the neighbours' counts)
    h->host_stale = true;
    if ((rc = slab_sync_host(h, false))) return rc;
    const uint32_t* hi = h->info_host;
    for (int k = 0; k < 8; k++) info[k] = (int32_t)hi[k];
    if (hi[5] & kSlabErrOwned)
        return fail(WC_ERR_CAPACITY, "slab holds %u particles, capacity %d", hi[0], h->cap);
    if (hi[5] & kSlabErrGhost)
        return fail(WC_ERR_CAPACITY, "halo layer of %u / %u particles exceeds slab_ghost_capacity %d",
                    hi[3], hi[4], h->Cg);
    if (hi[5] & kSlabErrHalo)
        return fail(WC_ERR_CAPACITY, "boundary layer of %d / %d particles exceeds the neighbours' "
                                     "slab_ghost_capacity %d", h->n_first, h->n_last, h->Cg);
    if (hi[5] & kSlabErrTimeout)
        return fail(WC_ERR_CUDA, "a neighbour's signal did not arrive (peer dead, or slab handles "
                                 "of one process stepped out of phase order)");
    return WC_OK;
}

int wc_slab_reorder(wc_handle* h) {
    WC_NEED_SLAB(h);
    if (h->step_no == 0) return fail(WC_ERR_INVALID, "wc_slab_reorder needs wc_slab_sort_count");
    // launch bounds by capacity: the virtual input [M | owned | M] and the owned region.  (The
    // ghost layers -- the neighbours' counts, scanned so the halo slices sit right before / after
    // the owned slice: [Cg - n_glow, Cg) and [Cg + n, Cg + n + n_ghigh) -- are queued inside.)
    return sort_reorder_phase(h, true, h->M + h->cap + h->M, h->cap);
}

int wc_slab_density(wc_handle* h, const wc_step_params* sp) {
    WC_NEED_SLAB(h);
    int rc = check_step_params(sp);
    if (rc) return rc;
    if (!h->sorted_valid) return fail(WC_ERR_INVALID, "wc_slab_density needs wc_slab_reorder");
    const bool simple = h->p.flags & WC_FLAG_SIMPLE_KERNELS;
    if (simple) {  // the cross-check kernels take their sizes from the host and do not signal
        if ((rc = slab_sync_host(h))) return rc;
        if ((rc = slab_sync_kernel(h, kSigHaloPos, -1, -1))) return rc;
    } else if ((rc = slab_pre_wait(h, kSigHaloPos))) {
        return rc;
    }
    if ((rc = run_density(h, *sp))) return rc;
    if (simple && h->peer_mode()) {
        if ((rc = peer_copy_halo(h))) return rc;
        if ((rc = slab_sync_kernel(h, -1, kSigHaloRho, kDoneDensity))) return rc;
    }
    return record(h, 4);
}

int wc_slab_update(wc_handle* h, float frame_dt, const wc_step_params* sp) {
    return slab_update_impl(h, frame_dt, sp, nullptr);
}

int wc_slab_step_peer(wc_handle* h, float frame_dt, const wc_step_params* sp, int32_t info[8]) {
    return slab_step_impl(h, frame_dt, sp, info, nullptr);
}

int wc_slab_step_peer_host(wc_handle* h, float frame_dt, const wc_step_params* sp,
                           const wc_particle* host_in, int32_t n, wc_particle* host_out,
                           int32_t out_capacity, int32_t info[8]) {
    if (!info) return fail(WC_ERR_INVALID, "info is NULL (the caller needs the new particle count)");
    WC_NEED_SLAB(h);
    int rc;
    if (host_in && (rc = wc_upload_particles(h, host_in, n))) return rc;
    if (!host_out) return fail(WC_ERR_INVALID, "host_out is NULL");
    if (out_capacity < h->cap)
        return fail(WC_ERR_CAPACITY, "host_out holds %d particles, the slab may own %d", out_capacity,
                    h->cap);
    // Page-locked destination: the update kernel writes the AoS records into it directly.
    float4* mapped = nullptr;
    if (!(h->p.flags & WC_FLAG_SIMPLE_KERNELS)) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, host_out) == cudaSuccess &&
            attr.type == cudaMemoryTypeHost && attr.devicePointer)
            mapped = static_cast<float4*>(attr.devicePointer);
        cudaGetLastError();  // an unregistered pointer is not an error here
    }
    if ((rc = slab_step_impl(h, frame_dt, sp, info, mapped))) return rc;  // syncs (info != NULL)
    if (!mapped) return wc_download_particles(h, 1, host_out);
    return WC_OK;
}

int wc_slab_ipc_export(wc_handle* h, wc_slab_ipc* out) {
    if (!out) return fail(WC_ERR_INVALID, "NULL argument");
    WC_NEED_SLAB(h);
    static_assert(sizeof(cudaIpcMemHandle_t) == WC_IPC_HANDLE_BYTES, "IPC handle size");
    void* ptrs[7] = {h->pos[1], h->vel[1], h->mig_in[0], h->mig_in[1],
                     h->lc_recv[0], h->lc_recv[1], h->sig};
    for (int k = 0; k < 7; k++) {
        cudaIpcMemHandle_t mh;
        WC_CUDA(cudaIpcGetMemHandle(&mh, ptrs[k]));
        std::memcpy(out->mem[k], &mh, sizeof(mh));
    }
    out->device = h->p.device;
    out->ghost_capacity = h->Cg;
    return WC_OK;
}

static int peer_check(wc_handle* h, int32_t direction) {
    if (direction != 0 && direction != 1) return fail(WC_ERR_INVALID, "direction must be 0 or 1");
    if (h->peer[direction].on) return fail(WC_ERR_INVALID, "neighbour %d is already attached", direction);
    if (h->step_no != 0) return fail(WC_ERR_INVALID, "attach neighbours before the first step");
    return WC_OK;
}

int wc_slab_peer_open(wc_handle* h, int32_t direction, const wc_slab_ipc* peer) {
    if (!peer) return fail(WC_ERR_INVALID, "NULL argument");
    WC_NEED_SLAB(h);
    int rc = peer_check(h, direction);
    if (rc) return rc;
    if (peer->ghost_capacity != h->Cg)
        return fail(WC_ERR_INVALID, "neighbour has another slab_ghost_capacity (%d vs %d)",
                    peer->ghost_capacity, h->Cg);
    wc_handle::Peer& P = h->peer[direction];
    // what this rank writes at the neighbour: its sorted buffers, the message buffers that
    // face this rank (index 1 - direction), its signals
    const int want[5] = {0, 1, 2 + (1 - direction), 4 + (1 - direction), 6};
    void* mapped[5] = {};
    for (int k = 0; k < 5; k++) {
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, peer->mem[want[k]], sizeof(mh));
        cudaError_t e = cudaIpcOpenMemHandle(&mapped[k], mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < k; q++) cudaIpcCloseMemHandle(mapped[q]);
            return fail(WC_ERR_CUDA, "cudaIpcOpenMemHandle (neighbour %d, buffer %d): %s", direction,
                        want[k], cudaGetErrorString(e));
        }
        P.ipc_base[k] = mapped[k];
    }
    if (peer->device == h->p.device) h->same_device_peer = true;
    P.pos1 = (float4*)mapped[0];
    P.vel1 = (float4*)mapped[1];
    P.mig_in = (float4*)mapped[2];
    P.lc_recv = (uint32_t*)mapped[3];
    P.sig = (uint32_t*)mapped[4];
    P.ipc = true;
    P.on = true;
    return WC_OK;
}

int wc_slab_peer_attach(wc_handle* h, int32_t direction, wc_handle* peer) {
    if (!peer) return fail(WC_ERR_INVALID, "NULL argument");
    WC_NEED_SLAB(h);
    if (!peer->slab) return fail(WC_ERR_INVALID, "the neighbour is not a slab handle");
    int rc = peer_check(h, direction);
    if (rc) return rc;
    if (peer->Cg != h->Cg) return fail(WC_ERR_INVALID, "neighbour has another slab_ghost_capacity");
    if (peer->p.device != h->p.device) {  // same process, other device: plain peer access
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->p.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(WC_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    wc_handle::Peer& P = h->peer[direction];
    if (peer->p.device == h->p.device) h->same_device_peer = true;
    P.pos1 = peer->pos[1];
    P.vel1 = peer->vel[1];
    P.mig_in = peer->mig_in[1 - direction];
    P.lc_recv = peer->lc_recv[1 - direction];
    P.sig = peer->sig;
    P.on = true;
    return WC_OK;
}

int wc_diagnose(wc_handle* h, int32_t which, float rest_density, wc_diagnostics* out) {
    if (!h || !out) return fail(WC_ERR_INVALID, "NULL argument");
    if (which != 1 && which != 2) return fail(WC_ERR_INVALID, "which must be 1 or 2");
    if (!(rest_density > 0.0f)) return fail(WC_ERR_INVALID, "rest_density must be positive");
    WC_CUDA(cudaSetDevice(h->p.device));
    { int rcs = slab_sync_host(h); if (rcs) return rcs; }
    if (!h->diag) {  // allocated on first use: most runs never ask
        WC_CUDA(cudaMalloc(&h->diag, (size_t)(kDiagMaxBlocks + 1) * sizeof(DiagPartial)));
        WC_CUDA(cudaMalloc(&h->diag_cells, 2 * sizeof(unsigned int)));
    }
    std::memset(out, 0, sizeof(*out));
    out->particles = h->n;
    out->max_cell_count = out->nonempty_cells = -1;
    DiagPartial r;
    std::memset(&r, 0, sizeof(r));
    unsigned int cells[2] = {0u, 0u};
    const bool have_cells = h->sorted_valid && !h->slab;
    if (h->n > 0) {
        const int first = which == 1 ? h->M : h->Cg;
        // a fixed slice per block, so the fold order depends on n alone
        const int blocks = (int)std::min<long long>(kDiagMaxBlocks, div_up(h->n, 4 * kDiagThreads));
        const int per_block = div_up(h->n, blocks);
        k_diag_particles<<<blocks, kDiagThreads, 0, h->stream>>>(
            h->pos[which - 1] + first, h->vel[which - 1] + first, h->n, per_block, h->p.size,
            1.0f / rest_density, h->diag + 1);
        WC_CHECK_LAUNCH(h);
        k_diag_fold<<<1, kDiagThreads, 0, h->stream>>>(h->diag + 1, blocks, h->diag);
        WC_CHECK_LAUNCH(h);
        WC_CUDA(cudaMemcpyAsync(&r, h->diag, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
    }
    if (have_cells) {
        WC_CUDA(cudaMemsetAsync(h->diag_cells, 0, sizeof(cells), h->stream));
        k_diag_cells<<<div_up(h->num_bins, 256), 256, 0, h->stream>>>(h->offsets, h->num_bins,
                                                                     h->diag_cells);
        WC_CHECK_LAUNCH(h);
        WC_CUDA(cudaMemcpyAsync(cells, h->diag_cells, sizeof(cells), cudaMemcpyDeviceToHost,
                                h->stream));
    }
    WC_CUDA(cudaStreamSynchronize(h->stream));
    const double m = (double)h->d.particle_mass, nv = (double)r.valid;
    out->invalid = (int64_t)r.invalid;
    out->out_of_box = (int64_t)r.out_of_box;
    out->at_speed_clamp = (int64_t)r.at_clamp;
    out->mass = m * nv;
    for (int a = 0; a < 3; a++) {
        out->momentum[a] = m * r.mom[a];
        out->centre_of_mass[a] = r.valid ? r.com[a] / nv : 0.0;
    }
    out->kinetic_energy = 0.5 * m * r.ke;
    if (r.valid) {
        out->max_speed = std::sqrt((double)r.vmax2);
        out->density_min = r.rho_min, out->density_max = r.rho_max;
        out->density_mean = r.rho_sum / nv;
        out->pressure_min = r.pres_min, out->pressure_max = r.pres_max;
        out->pressure_mean = r.pres_sum / nv;
    }
    for (int k = 0; k < WC_DIAG_HIST_BINS; k++) out->density_hist[k] = (int64_t)r.hist[k];
    if (have_cells) {
        out->max_cell_count = cells[0];
        out->nonempty_cells = cells[1];
    }
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
