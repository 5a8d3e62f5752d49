// wc_sort.cuh -- cell-hash binning and the STABLE counting sort.
//
// Replaces the reference's three sort dispatches (src/core/Sort.cpp:254-267):
//   count.comp:25-36      -> k_hash_count   (warp-aggregated atomics, also records the
//                                            arrival rank so no second atomic pass runs)
//   linearScan.comp:15-28 -> k_scan         (single-pass decoupled look-back scan instead of
//                                            the 1-thread serial loop, Sort.cpp:181-184)
//   reorder.comp:32-45    -> k_scatter_ids + k_reorder (sort.comp:32-45's ID scatter followed
//                                            by an in-cell rank fix-up that makes the order
//                                            the canonical ascending-input-index one, Q1)
#pragma once

#include "wc_common.cuh"

namespace wc {

// ---------------------------------------------------------------------------------------
// AoS (32-byte struct Particle, src/core/util.h:29-35) <-> SoA float4 pairs.
__global__ void k_aos_to_soa(const float4* __restrict__ aos, int n, float4* __restrict__ pos_rho,
                             float4* __restrict__ vel_pres) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos_rho[i] = aos[2 * (size_t)i];
    vel_pres[i] = aos[2 * (size_t)i + 1];
}

__global__ void k_soa_to_aos(const float4* __restrict__ pos_rho,
                             const float4* __restrict__ vel_pres, int n,
                             float4* __restrict__ aos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    aos[2 * (size_t)i] = pos_rho[i];
    aos[2 * (size_t)i + 1] = vel_pres[i];
}

// ---------------------------------------------------------------------------------------
// count.comp:25-36.  One thread per particle; lanes of a warp that hit the same cell are
// merged with __match_any_sync so one atomicAdd serves the whole group (the input of
// step k+1 is the cell-sorted output of step k, so groups are long).  rank = arrival
// order within the cell; it is NOT the stable rank (k_reorder fixes that up).
__global__ void __launch_bounds__(256)
k_hash_count(const float4* __restrict__ pos_rho, int n, float bin, int G,
             uint32_t* __restrict__ cell_ids, uint32_t* __restrict__ ranks,
             uint32_t* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    uint32_t c = 0xFFFFFFFFu;
    if (active) {
        const float4 p = pos_rho[i];
        c = cell_index(p.x, p.y, p.z, bin, G);
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned group = __match_any_sync(0xffffffffu, c);
    if (active) {
        const int leader = __ffs(group) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(&counts[c], (uint32_t)__popc(group));
        base = __shfl_sync(group, base, leader);
        cell_ids[i] = c;
        ranks[i] = base + (uint32_t)__popc(group & ((1u << lane) - 1u));
    }
}

// ---------------------------------------------------------------------------------------
// linearScan.comp:15-28 as a single-pass chained scan with decoupled look-back.
// status[tile] packs {flag:2 | value:32} in one 64-bit word so flag and value are
// published atomically.  status[] and *tile_counter must be zero at launch (they live in
// the same per-step cleared arena as counts[]).
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr unsigned long long kFlagAgg = 1ull << 32;
constexpr unsigned long long kFlagIncl = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

// Decoupled look-back of one tile (call with all 32 lanes of ONE warp): publishes the tile's
// aggregate, folds the predecessors' status words and returns the tile's exclusive prefix;
// finally publishes the inclusive prefix.  status[] must be zero at launch.
__device__ __forceinline__ uint32_t lookback_exclusive_prefix(unsigned long long* status, int tile,
                                                              uint32_t aggregate, int lane) {
    uint32_t prefix = 0;
    if (tile > 0) {
        if (lane == 0) st_status(&status[tile], kFlagAgg | aggregate);
        int look = tile - 1;
        while (true) {
            const int idx = look - lane;
            unsigned long long s = kFlagIncl;  // virtual tile -1: inclusive prefix 0
            if (idx >= 0) {
                do {
                    s = ld_status(&status[idx]);
                } while ((s >> 32) == 0ull);
            }
            const unsigned incl_mask = __ballot_sync(0xffffffffu, (s >> 32) == 2ull);
            uint32_t val = (uint32_t)s;
            if (incl_mask) {
                const int first = __ffs(incl_mask) - 1;
                if (lane > first) val = 0;
                prefix += warp_sum(val);
                break;
            }
            prefix += warp_sum(val);
            look -= 32;
        }
    }
    if (lane == 0) st_status(&status[tile], kFlagIncl | (unsigned long long)(prefix + aggregate));
    return prefix;
}

// Block-wide exclusive scan of one value per thread (kScanThreads threads); returns the
// thread's exclusive prefix, *total = the block's sum.  s_warp: 32 words of shared memory.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0u, winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
    }
    *total = __shfl_sync(0xffffffffu, winc, 31);
    return __shfl_sync(0xffffffffu, winc - w, warp) + (incl - v);
}

// Writes out[i] = out_base + sum(in[0..i)) for i < n and out[n] = out_base + sum(in[0..n)).
__global__ void __launch_bounds__(kScanThreads)
k_scan(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n,
       unsigned long long* __restrict__ status, unsigned int* __restrict__ tile_counter,
       uint32_t out_base) {
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_warp[kScanThreads / kWarp];
    __shared__ uint32_t s_prefix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int base = tile * kScanTile + tid * kScanItems;

    uint32_t v[kScanItems];
    // base is a multiple of 4; in[] itself is only 16-byte aligned for the whole-grid table
    // (slab mode scans from counts + G*G, any alignment), so the vector path checks it.
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    if (vec_ok && base + kScanItems <= n) {
        const uint4 q = *reinterpret_cast<const uint4*>(in + base);
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) v[k] = (base + k < n) ? in[base + k] : 0u;
    }
    const uint32_t tsum = v[0] + v[1] + v[2] + v[3];

    uint32_t aggregate;
    const uint32_t excl = block_exclusive_scan(tsum, s_warp, &aggregate);
    if (warp == 0) {
        const uint32_t prefix = lookback_exclusive_prefix(status, tile, aggregate, lane);
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();

    uint32_t run = out_base + s_prefix + excl;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
        if (base + k == n - 1) out[n] = run;
    }
    if (n == 0 && tile == 0 && tid == 0) out[0] = out_base;
}

// ---------------------------------------------------------------------------------------
// sort.comp:41-44: sorted[offsets[cell] + localOffset] = particleID, with the arrival rank
// recorded by k_hash_count standing in for the second atomicAdd pass.
// `base` is the index the offsets table assigns to the first sorted slot (0 unless slab
// mode); cell id 0xFFFFFFFF marks an input slot that does not take part (slab mode).
// slab mode: the input is the virtual array [M migrant slots | owned | M] whose owned count is
// in the slab record (n is then only the launch bound).
__global__ void __launch_bounds__(256)
k_scatter_ids(const uint32_t* __restrict__ cell_ids, const uint32_t* __restrict__ ranks,
              const uint32_t* __restrict__ offsets, int n, uint32_t* __restrict__ ids,
              uint32_t base, SlabRef slab, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (slab.dyn) {
        if (slab.dyn->errors) return;
        n = M + (int)slab.dyn->n_in_old + M;
    }
    if (i >= n) return;
    const uint32_t c = cell_ids[i];
    if (c == 0xFFFFFFFFu) return;
    ids[offsets[c] - base + ranks[i]] = (uint32_t)i;
}

// reorder.comp:32-45 made stable.  One thread per slot j of the arrival-ordered ID list:
// it owns particle id = ids[j], finds its cell's slice [beg, end) and counts the IDs in the
// slice that are smaller: that is the stable rank (what the serial oracle loop produces).
// Threads of one cell walk the same slice in lockstep, so the loads are warp broadcasts.
// The payload moves as two float4 (SoA), and the permutation is kept (sort.comp's output).
// The in-cell count is quadratic in the cell's occupancy, which is fine at the reference's
// ~20 particles per cell; a cell holding more than kBigCell particles (all particles piled
// into a few cells, NaN positions collapsing into cell 0, gridRes of 1..few) is only
// registered here and handled by reorder_big_cells (k_finish_sort).
constexpr int kBigCell = 256;
constexpr int kBigThreads = 1024;
constexpr int kBigBlocks = 148;  // blocks stride over the registered cells

struct ReorderIO {
    const float4* pos_in;
    const float4* vel_in;
    float4* pos_out;
    float4* vel_out;
    uint32_t* perm;
};

__device__ __forceinline__ void reorder_store(const ReorderIO& io, const PeerHalo& peer, uint32_t dst,
                                              uint32_t id, float4 p, float4 v) {
    io.pos_out[dst] = p;
    io.vel_out[dst] = v;
    io.perm[dst] = id;
    // halo positions (+ velocities, which do not change before the update) straight into the
    // neighbours' ghost slots
    if (peer.pos[0] && dst < peer.n_first) {
        peer.pos[0][peer.dst[0] + dst] = p;
        peer.vel[0][peer.dst[0] + dst] = v;
    }
    if (peer.pos[1] && dst >= peer.hi_begin) {
        peer.pos[1][peer.dst[1] + (dst - peer.hi_begin)] = p;
        peer.vel[1][peer.dst[1] + (dst - peer.hi_begin)] = v;
    }
}

__global__ void __launch_bounds__(256)
k_reorder(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ offsets, int n, float bin,
          int G, int zbase, uint32_t base, ReorderIO io, SlabRef slab,
          uint32_t* __restrict__ big_cells, uint32_t* __restrict__ big_count, uint32_t big_cap) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (slab.dyn) {  // slab mode: n is the launch bound, the owned count is in the slab record
        if (slab.dyn->errors) return;
        n = (int)slab.dyn->n;
    }
    if (j >= n) return;
    const PeerHalo peer = peer_halo_of(slab);
    const uint32_t id = ids[j];
    const float4 p = io.pos_in[id];
    const float4 v = io.vel_in[id];  // in flight under the rank loop
    const uint32_t c = cell_index(p.x, p.y, p.z, bin, G, zbase);
    const uint32_t beg = offsets[c] - base, end = offsets[c + 1] - base;
    if (end - beg > (uint32_t)kBigCell) {
        if ((uint32_t)j == beg) {  // one registration per big cell
            const uint32_t e = atomicAdd(big_count, 1u);
            if (e < big_cap) big_cells[e] = c;
        }
        return;
    }
    uint32_t rank = 0;
    for (uint32_t k = beg; k < end; k++) rank += (ids[k] < id) ? 1u : 0u;
    reorder_store(io, peer, beg + rank, id, p, v);
}

// Cells above kBigCell (second part of k_finish_sort): one block per registered cell sorts
// the cell's arrival-ordered ID slice ascending (= the stable order) with an LSD radix sort,
// 8 bits per pass, ping-ponging between ids[] and scratch[] inside the slice; the last pass
// moves the payload instead of storing the ID.  Linear in the cell's occupancy per pass, so one
// cell holding every particle costs O(n), not O(n^2).
// Each pass walks the slice in tiles of kBigThreads IDs, in order: a tile is ranked stably
// (match_any inside a warp, a per-digit prefix over the tile's warps) and scattered behind the
// running digit offsets.  The digit histogram of pass p + 1 is taken while pass p scatters, and
// a tile's IDs are loaded one tile ahead.
__device__ __forceinline__ void
reorder_big_cells(uint32_t* __restrict__ ids, uint32_t* __restrict__ scratch,
                  const uint32_t* __restrict__ offsets, uint32_t base, const ReorderIO& io,
                  const PeerHalo& peer, const uint32_t* __restrict__ big_cells,
                  const uint32_t* __restrict__ big_count, uint32_t big_cap, int passes) {
    __shared__ uint32_t s_bin[256];                  // running start of every digit's output range
    __shared__ uint32_t s_next[256];                 // histogram of the next pass's digit
    __shared__ uint32_t s_tot[256];
    __shared__ uint16_t s_warp[kBigThreads / 32][256];  // per-warp digit counts of one tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nbig = min(*big_count, big_cap);
    for (uint32_t e = blockIdx.x; e < nbig; e += gridDim.x) {
        const uint32_t c = big_cells[e];
        const uint32_t beg = offsets[c] - base, k = offsets[c + 1] - base - beg;
        uint32_t* src = ids + beg;
        uint32_t* dst = scratch + beg;
        if (tid < 256) s_next[tid] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < k; i += kBigThreads) atomicAdd(&s_next[src[i] & 255u], 1u);
        __syncthreads();
        for (int pass = 0; pass < passes; pass++) {
            const int shift = 8 * pass;
            const bool last = pass + 1 == passes;
            if (warp == 0) {  // exclusive scan of the 256 digit totals: 8 per lane
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = s_next[lane * 8 + q], sum += v[q];
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                uint32_t run = incl - sum;
#pragma unroll
                for (int q = 0; q < 8; q++) s_bin[lane * 8 + q] = run, run += v[q];
            }
            __syncthreads();
            if (tid < 256) s_next[tid] = 0;
            uint32_t id_next = tid < k ? src[tid] : 0u;
            __syncthreads();
            for (uint32_t t0 = 0; t0 < k; t0 += kBigThreads) {
                const uint32_t i = t0 + tid;
                const bool valid = i < k;
                const uint32_t id = id_next;
                if (i + kBigThreads < k) id_next = src[i + kBigThreads];  // one tile ahead
                const uint32_t d = valid ? (id >> shift) & 255u : 0xFFFFFFFFu;
#pragma unroll
                for (int q = 0; q < 8; q++) s_warp[warp][lane * 8 + q] = 0;
                __syncwarp();
                const unsigned same = __match_any_sync(0xffffffffu, d);
                const uint32_t in_warp = (uint32_t)__popc(same & ((1u << lane) - 1u));
                if (valid && in_warp == 0) s_warp[warp][d] = (uint16_t)__popc(same);
                if (valid && !last) atomicAdd(&s_next[(id >> (shift + 8)) & 255u], 1u);
                __syncthreads();
                if (tid < 256) {  // exclusive prefix over the tile's warps, per digit
                    uint32_t run = 0;
#pragma unroll 8
                    for (int w = 0; w < kBigThreads / 32; w++) {
                        const uint32_t x = s_warp[w][tid];
                        s_warp[w][tid] = (uint16_t)run;
                        run += x;
                    }
                    s_tot[tid] = run;
                }
                __syncthreads();
                if (valid) {
                    const uint32_t at = s_bin[d] + s_warp[warp][d] + in_warp;
                    if (last) reorder_store(io, peer, beg + at, id, io.pos_in[id], io.vel_in[id]);
                    else dst[at] = id;
                }
                __syncthreads();
                if (tid < 256) s_bin[tid] += s_tot[tid];
            }
            __syncthreads();  // this pass's global stores are read by the next one
            uint32_t* t = src;
            src = dst, dst = t;
        }
    }
}

}  // namespace wc
