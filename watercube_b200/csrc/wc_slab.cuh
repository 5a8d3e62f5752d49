// wc_slab.cuh -- device pieces of the z-slab decomposition (multi-GPU).
//
// The reference is single-GPU; this is new work.  The cell index is z-major
// (count.comp:33), so a rank that owns z-layers [z_begin, z_end) owns a contiguous range of
// bins and, after the sort, a contiguous slice of the particle array; its first / last
// layer (the halo the neighbours need) are contiguous sub-slices, so halo sends need no
// packing.  Layout on one rank (slab-local table of z_end - z_begin + 2 layers):
//
//   buffer 1 (input):   [ M slots: migrants from below | owned, previous order | M: from above ]
//   buffer 2 (sorted):  [ ... ghost-low layer ][ owned, cell-sorted ][ ghost-high layer ... ]
//                                             ^ index Cg
// Concatenating the ranks' inputs in z order reproduces the single-GPU input order for every
// cell (particles that arrive from the rank below come first, from above last), so the stable
// sort -- and with it every result -- is bit-identical to the undecomposed run.
//
// A step is eleven kernels (+ one memset), none of whose launch parameters depends on a value
// the step itself computes: the counts live in the device record SlabDyn (wc_common.cuh), the
// grids are sized by capacity, and with attached neighbours (wc_slab_peer_*) every message is
// stored into the neighbour's memory by the kernel that produces it, with the wait for a
// message in the consuming kernel's first instructions and the signal in the producing
// kernel's last block.  The host therefore never waits inside a step.
#pragma once

#include "wc_common.cuh"
#include "wc_sort.cuh"

namespace wc {

constexpr int kMigHeaderFloat4 = 2;  // 32-byte message header: [count, 7 x pad]
constexpr int kLcHeader = 2;         // layer-count message header: [n_layer, n_owned]

// First kernel of a step: (waits for last step's migrant messages and) unpacks them -- header
// + AoS payload -> the SoA slots before / after the owned region of buffer 1.
__global__ void __launch_bounds__(256)
k_slab_begin(const float4* __restrict__ msg_below, const float4* __restrict__ msg_above, int M,
             float4* __restrict__ pos, float4* __restrict__ vel, SlabRef slab) {
    slab_block_wait(slab);
    SlabDyn* dyn = slab.dyn;
    const uint32_t c0 = min(reinterpret_cast<const uint32_t*>(msg_below)[0], (uint32_t)M);
    const uint32_t c1 = min(reinterpret_cast<const uint32_t*>(msg_above)[0], (uint32_t)M);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) dyn->m_in[0] = c0, dyn->m_in[1] = c1;
    if (i >= 2 * M) return;
    const bool above = i >= M;
    const uint32_t k = above ? (uint32_t)(i - M) : (uint32_t)i;
    if (k >= (above ? c1 : c0)) return;
    const float4* msg = above ? msg_above : msg_below;
    const size_t dst = above ? (size_t)M + dyn->n_in_old + k : (size_t)k;
    pos[dst] = msg[kMigHeaderFloat4 + 2 * (size_t)k];
    vel[dst] = msg[kMigHeaderFloat4 + 2 * (size_t)k + 1];
}

// count.comp:25-36 over the virtual input [from-below | owned | from-above].  A slot takes
// part when it holds a received migrant, or an owned particle that is still inside the slab
// (owned particles that left were sent to the neighbour at the end of the previous step).
// migrants_only: the owned slots were hashed and counted by the previous step's update pass
// (PreHash); the 2 M threads cover the two migrant windows.
__global__ void __launch_bounds__(256)
k_hash_count_slab(const float4* __restrict__ pos, int M, SlabDyn* __restrict__ dyn, float bin,
                  int G, int z_begin, int z_end, uint32_t* __restrict__ cell_ids,
                  uint32_t* __restrict__ ranks, uint32_t* __restrict__ counts, int migrants_only) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_old = (int)dyn->n_in_old;
    bool active = i < M + n_old + M;
    if (migrants_only) {
        active = i < 2 * M;
        if (i >= M) i += n_old;  // the window behind the owned region
    }
    uint32_t c = 0xFFFFFFFFu;
    if (active) {
        const bool own = i >= M && i < M + n_old;
        const bool valid =
            own || (i < M ? (uint32_t)i < dyn->m_in[0] : (uint32_t)(i - M - n_old) < dyn->m_in[1]);
        if (valid) {
            const float4 p = pos[i];
            const int cz = cell_coord(p.z, bin, G);
            if (cz >= z_begin && cz < z_end) {
                c = cell_index(p.x, p.y, p.z, bin, G, z_begin - 1);
            } else if (!own) {
                atomicOr(&dyn->errors, kSlabErrStray);  // a migrant that moved more than one layer
            }
        }
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned group = __match_any_sync(0xffffffffu, c);
    if (active) {
        cell_ids[i] = c;
        if (c != 0xFFFFFFFFu) {
            const int leader = __ffs(group) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&counts[c], (uint32_t)__popc(group));
            base = __shfl_sync(group, base, leader);
            ranks[i] = base + (uint32_t)__popc(group & ((1u << lane) - 1u));
        }
    }
}

// After the owned-layer scan: counts of the slab and its boundary layers into the slab record,
// and the layer-count messages [n_layer, n_owned, counts of the layer's G*G cells] -- into the
// local send buffers (host-driven transports) and straight into the attached neighbours'
// receive buffers, whose "layer counts" flags the last block raises.
__global__ void __launch_bounds__(256)
k_slab_info(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, int G2,
            int Lz, uint32_t* __restrict__ lc_down, uint32_t* __restrict__ lc_up,
            uint32_t* __restrict__ peer_down, uint32_t* __restrict__ peer_up, SlabRef slab) {
    const uint32_t Cg = slab.Cg;
    const uint32_t n_own = offsets[(size_t)(Lz - 1) * G2] - Cg;
    const uint32_t n_first = offsets[(size_t)2 * G2] - Cg;
    const uint32_t n_last = n_own - (offsets[(size_t)(Lz - 2) * G2] - Cg);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // peer_down / peer_up: the neighbours' receive buffers (mapped peer memory) or null
    if (i == 0) {
        SlabDyn* dyn = slab.dyn;
        dyn->n = n_own, dyn->n_first = n_first, dyn->n_last = n_last;
        uint32_t err = 0u;
        if (n_own > slab.cap) err |= kSlabErrOwned;
        if ((peer_down && n_first > Cg) || (peer_up && n_last > Cg)) err |= kSlabErrHalo;
        if (err) atomicOr(&dyn->errors, err);
        lc_down[0] = n_first, lc_down[1] = n_own;
        lc_up[0] = n_last, lc_up[1] = n_own;
        if (peer_down) peer_down[0] = n_first, peer_down[1] = n_own;
        if (peer_up) peer_up[0] = n_last, peer_up[1] = n_own;
    }
    if (i < G2) {
        const uint32_t cd = counts[(size_t)G2 + i], cu = counts[(size_t)(Lz - 2) * G2 + i];
        lc_down[kLcHeader + i] = cd;
        lc_up[kLcHeader + i] = cu;
        if (peer_down) peer_down[kLcHeader + i] = cd;
        if (peer_up) peer_up[kLcHeader + i] = cu;
    }
    slab_grid_signal(slab, peer_down != nullptr || peer_up != nullptr, gridDim.x);
}

// The two ghost layers of the table, one block each: (waits for the neighbour's layer-count
// message,) copies its counts into the table and scans them so the halo slices sit right
// before / after the owned slice of buffer 2: [Cg - n_glow, Cg) and [Cg + n, Cg + n + n_ghigh).
// Also completes the slab record and mirrors it into page-locked host memory (info_host), for
// whoever wants this step's counts later (wc_slab_sync_info; nobody has to).
// info_host: [n_own, n_first, n_last, ghost below, ghost above, errors, migrants in (2),
//             neighbours' n_own (2)]
__global__ void __launch_bounds__(kScanThreads)
k_ghost_tables(const uint32_t* __restrict__ lc_below, const uint32_t* __restrict__ lc_above, int G2,
               int Lz, uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
               uint32_t* __restrict__ info_host, SlabRef slab) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_run;
    const int dir = blockIdx.x;  // 0: the layer below the slab, 1: the layer above
    // only the flag of this block's direction matters
    SlabRef w = slab;
    w.wait[0] = dir == 0 ? slab.wait[0] : slab.wait[1];
    w.wait[1] = nullptr;
    slab_block_wait(w);
    SlabDyn* dyn = slab.dyn;
    const uint32_t* lc = dir == 0 ? lc_below : lc_above;
    const uint32_t n_ghost = lc[0], n_peer = lc[1];
    const uint32_t n = dyn->n;
    const size_t layer = dir == 0 ? 0 : (size_t)(Lz - 1) * G2;
    // a ghost layer beyond the capacity kills the step (see slab_dead); the scan below then
    // only has to stay inside the table, which it does for any counts
    const uint32_t base = dir == 0 ? slab.Cg - min(n_ghost, slab.Cg) : slab.Cg + n;
    if (threadIdx.x == 0) {
        s_run = 0u;
        if (dir == 0) dyn->n_glow = n_ghost, dyn->peer_n[0] = n_peer;
        else dyn->n_ghigh = n_ghost, dyn->peer_n[1] = n_peer;
        if (n_ghost > slab.Cg) atomicOr(&dyn->errors, kSlabErrGhost);
    }
    __syncthreads();
    for (int t0 = 0; t0 < G2; t0 += kScanTile) {
        const int b = t0 + (int)threadIdx.x * kScanItems;
        uint32_t v[kScanItems], tsum = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            v[k] = (b + k < G2) ? lc[kLcHeader + b + k] : 0u;
            if (b + k < G2) counts[layer + b + k] = v[k];
            tsum += v[k];
        }
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(tsum, s_warp, &total);
        uint32_t run = base + s_run + excl;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            if (b + k < G2) offsets[layer + b + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (dir == 1) offsets[layer + G2] = base + s_run;  // the table's closing sentinel
        if (dir == 0) {
            info_host[0] = n, info_host[1] = dyn->n_first, info_host[2] = dyn->n_last;
            info_host[3] = n_ghost, info_host[6] = dyn->m_in[0], info_host[7] = dyn->m_in[1];
            info_host[8] = n_peer;
        } else {
            info_host[4] = n_ghost, info_host[9] = n_peer;
        }
        __threadfence();
        // errors last, by whichever block finishes second, so it covers both directions
        if (atomicAdd(slab.done, 1u) == 1u) {
            __threadfence();
            info_host[5] = dyn->errors;
        }
    }
}

// Last kernel of a step: the particles of the first / last owned layer whose new z-layer left
// the slab (only those layers can lose any: |v| dt < binSize, and they are the head / tail of
// the sorted order), compacted in order -- a chained scan with decoupled look-back per
// direction -- into the outgoing message: the local buffer (host-driven transports) and the
// attached neighbour's receive buffer, whose "migrants" flag the last block raises.  The same
// block closes the step: the output's owned count becomes the next step's input count.
// grid = (tiles, 2 directions).  dir 0: below z_begin (to rank - 1); dir 1: at or above z_end.
__global__ void __launch_bounds__(kScanThreads)
k_migrants(const float4* __restrict__ pos, const float4* __restrict__ vel, float bin, int G,
           int z_begin, int z_end, int M, float4* __restrict__ msg_down, float4* __restrict__ msg_up,
           float4* __restrict__ peer_down, float4* __restrict__ peer_up,
           unsigned long long* __restrict__ status_down, unsigned long long* __restrict__ status_up,
           unsigned int* __restrict__ ticket_down, unsigned int* __restrict__ ticket_up,
           SlabRef slab) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_prefix;
    __shared__ unsigned int s_tile;
    SlabDyn* dyn = slab.dyn;
    const int dir = blockIdx.y, tid = threadIdx.x;
    // tiles are handed out in arrival order, so a tile never waits for one that is not running
    if (tid == 0) s_tile = atomicAdd(dir == 0 ? ticket_down : ticket_up, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const bool dead = slab_dead(slab);
    const uint32_t n = dyn->n;
    const uint32_t n_layer = dead ? 0u : (dir == 0 ? dyn->n_first : dyn->n_last);
    const uint32_t start = dir == 0 ? 0u : n - n_layer;
    float4* msg = dir == 0 ? msg_down : msg_up;
    float4* peer_msg = dir == 0 ? peer_down : peer_up;
    const int tiles_needed = (int)((n_layer + kScanTile - 1) / kScanTile);
    if (tile < max(tiles_needed, 1)) {  // tiles are consecutive from 0: the look-back chain is whole
        const uint32_t b = (uint32_t)tile * kScanTile + (uint32_t)tid * kScanItems;
        bool flag[kScanItems];
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            flag[k] = false;
            if (b + k < n_layer) {
                const int cz = cell_coord(pos[start + b + k].z, bin, G);
                flag[k] = dir == 0 ? cz < z_begin : cz >= z_end;
            }
            cnt += flag[k] ? 1u : 0u;
        }
        uint32_t aggregate;
        const uint32_t excl = block_exclusive_scan(cnt, s_warp, &aggregate);
        unsigned long long* status = dir == 0 ? status_down : status_up;
        if (tid < 32) {
            const uint32_t prefix = lookback_exclusive_prefix(status, tile, aggregate, tid);
            if (tid == 0) s_prefix = prefix;
        }
        __syncthreads();
        uint32_t slot = s_prefix + excl;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            if (!flag[k]) continue;
            if (slot < (uint32_t)M) {
                const float4 p = pos[start + b + k], v = vel[start + b + k];
                msg[kMigHeaderFloat4 + 2 * (size_t)slot] = p;
                msg[kMigHeaderFloat4 + 2 * (size_t)slot + 1] = v;
                if (peer_msg) {
                    peer_msg[kMigHeaderFloat4 + 2 * (size_t)slot] = p;
                    peer_msg[kMigHeaderFloat4 + 2 * (size_t)slot + 1] = v;
                }
            }
            slot++;
        }
        if (tile == max(tiles_needed, 1) - 1 && tid == kScanThreads - 1) {  // the direction's total
            const uint32_t total = s_prefix + excl + cnt;
            reinterpret_cast<uint32_t*>(msg)[0] = min(total, (uint32_t)M);
            if (peer_msg) reinterpret_cast<uint32_t*>(peer_msg)[0] = min(total, (uint32_t)M);
            if (total > (uint32_t)M) atomicOr(&dyn->errors, kSlabErrMigrants);
        }
    }
    // close the step (one thread of the whole grid), then signal
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        const uint32_t ticket = atomicAdd(slab.done, 1u);
        if (ticket == gridDim.x * gridDim.y - 1u) {
            __threadfence_system();
            dyn->n_in_old = dead ? dyn->n_in_old : n;
            if (slab.raise[0])
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(slab.raise[0]), "r"(slab.step_no) : "memory");
            if (slab.raise[1])
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(slab.raise[1]), "r"(slab.step_no) : "memory");
        }
    }
}

// Stand-alone wait and / or signal (one block) for the simple cross-check kernels, which carry
// neither.
__global__ void k_slab_sync(SlabRef slab) {
    slab_block_wait(slab);
    slab_grid_signal(slab, false, 1u);
}

}  // namespace wc
