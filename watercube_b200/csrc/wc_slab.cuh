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
#pragma once

#include "wc_common.cuh"

namespace wc {

constexpr int kMigHeaderFloat4 = 2;  // 32-byte message header: [count, 7 x pad]
constexpr int kLcHeader = 2;         // layer-count message header: [n_layer, n_owned]

// Incoming migrant message (header + AoS payload) -> SoA slots of buffer 1.
__global__ void k_unpack_migrants(const float4* __restrict__ msg, int cap,
                                  float4* __restrict__ pos_dst, float4* __restrict__ vel_dst,
                                  uint32_t* __restrict__ count_out) {
    const uint32_t count = min(reinterpret_cast<const uint32_t*>(msg)[0], (uint32_t)cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count_out = count;
    if ((uint32_t)i >= count) return;
    pos_dst[i] = msg[kMigHeaderFloat4 + 2 * (size_t)i];
    vel_dst[i] = msg[kMigHeaderFloat4 + 2 * (size_t)i + 1];
}

// count.comp:25-36 over the virtual input [from-below | owned | from-above].  A slot takes
// part when it holds a received migrant, or an owned particle that is still inside the slab
// (owned particles that left were sent to the neighbour at the end of the previous step).
__global__ void __launch_bounds__(256)
k_hash_count_slab(const float4* __restrict__ pos, int total, int M, int n_old,
                  const uint32_t* __restrict__ m_in, float bin, int G, int z_begin, int z_end,
                  uint32_t* __restrict__ cell_ids, uint32_t* __restrict__ ranks,
                  uint32_t* __restrict__ counts, uint32_t* __restrict__ errors) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < total;
    uint32_t c = 0xFFFFFFFFu;
    if (active) {
        const bool own = i >= M && i < M + n_old;
        const bool valid = own || (i < M ? (uint32_t)i < m_in[0] : (uint32_t)(i - M - n_old) < m_in[1]);
        if (valid) {
            const float4 p = pos[i];
            const int cz = cell_coord(p.z, bin, G);
            if (cz >= z_begin && cz < z_end) {
                c = cell_index(p.x, p.y, p.z, bin, G, z_begin - 1);
            } else if (!own) {
                atomicAdd(errors, 1u);  // a migrant that is not ours: moved more than one layer
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

// After the owned-layer scan: counts of the slab and its boundary layers, and the
// layer-count messages [n_layer, n_owned, counts of the layer's G*G cells] for the neighbours.
__global__ void k_slab_info(const uint32_t* __restrict__ counts,
                            const uint32_t* __restrict__ offsets, int G2, int Lz, uint32_t Cg,
                            uint32_t* __restrict__ info, uint32_t* __restrict__ lc_down,
                            uint32_t* __restrict__ lc_up, uint32_t* __restrict__ peer_down,
                            uint32_t* __restrict__ peer_up) {
    const uint32_t n_own = offsets[(size_t)(Lz - 1) * G2] - Cg;
    const uint32_t n_first = offsets[(size_t)2 * G2] - Cg;
    const uint32_t n_last = n_own - (offsets[(size_t)(Lz - 2) * G2] - Cg);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // peer_down / peer_up: the neighbours' receive buffers (mapped peer memory) or null
    if (i == 0) {
        info[0] = n_own, info[1] = n_first, info[2] = n_last;
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
}

// Everything wc_slab_sync_info hands to the host, gathered into one 10-word record that is
// stored straight into page-locked host memory (one kernel instead of seven small copies in
// the only host round trip of the step):
// [n_own, n_first, n_last, ghost below, ghost above, errors, migrants in (2), neighbours' n_own (2)]
__global__ void k_collect_info(const uint32_t* __restrict__ info, const uint32_t* __restrict__ lc_below,
                               const uint32_t* __restrict__ lc_above,
                               const uint32_t* __restrict__ errors,
                               const uint32_t* __restrict__ m_in, uint32_t* __restrict__ host_out) {
    host_out[0] = info[0], host_out[1] = info[1], host_out[2] = info[2];
    host_out[3] = lc_below[0], host_out[4] = lc_above[0];
    host_out[5] = *errors;
    host_out[6] = m_in[0], host_out[7] = m_in[1];
    host_out[8] = lc_below[1], host_out[9] = lc_above[1];
}

// Received layer counts -> the table's ghost layers.
__global__ void k_install_ghost_counts(const uint32_t* __restrict__ lc_below,
                                       const uint32_t* __restrict__ lc_above, int G2, int Lz,
                                       uint32_t* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G2) return;
    counts[i] = lc_below[kLcHeader + i];
    counts[(size_t)(Lz - 1) * G2 + i] = lc_above[kLcHeader + i];
}

// End of step: particles of the first / last owned layer whose new z-layer left the slab.
// dir 0: below z_begin (go to rank - 1); dir 1: at or above z_end (go to rank + 1).
__global__ void k_flag_migrants(const float4* __restrict__ pos, int n, float bin, int G,
                                int z_limit, int dir, uint32_t* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cz = cell_coord(pos[i].z, bin, G);
    flags[i] = (dir == 0 ? cz < z_limit : cz >= z_limit) ? 1u : 0u;
}

// Stable compaction (order preserved: see the file comment) into the outgoing message.
__global__ void k_pack_migrants(const float4* __restrict__ pos, const float4* __restrict__ vel,
                                int n, const uint32_t* __restrict__ flags,
                                const uint32_t* __restrict__ slot, int cap,
                                float4* __restrict__ msg, float4* __restrict__ peer_msg,
                                uint32_t* __restrict__ errors) {
    // peer_msg: the neighbour's mig_in (mapped peer memory) or null; gets the same message
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        const uint32_t total = slot[n];
        reinterpret_cast<uint32_t*>(msg)[0] = min(total, (uint32_t)cap);
        if (peer_msg) reinterpret_cast<uint32_t*>(peer_msg)[0] = min(total, (uint32_t)cap);
        if (total > (uint32_t)cap) atomicAdd(errors, total - (uint32_t)cap);
    }
    if (i >= n || !flags[i]) return;
    const uint32_t s = slot[i];
    if (s >= (uint32_t)cap) return;
    const float4 p = pos[i], v = vel[i];
    msg[kMigHeaderFloat4 + 2 * (size_t)s] = p;
    msg[kMigHeaderFloat4 + 2 * (size_t)s + 1] = v;
    if (peer_msg) {
        peer_msg[kMigHeaderFloat4 + 2 * (size_t)s] = p;
        peer_msg[kMigHeaderFloat4 + 2 * (size_t)s + 1] = v;
    }
}

__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }

// ---- peer-memory exchange (wc_slab_peer_*): put-with-signal over NVLink -----------------------
// The sender copies its data into the neighbour's buffers (mapped peer memory) on its own
// stream and then raises a monotonic step counter in the neighbour's signal array; the
// neighbour's stream holds a one-thread kernel that spins on its local counter before the
// consumer kernels run.  No host round trip, no collective library on the data path.
__global__ void k_signal(uint32_t* peer_flag, uint32_t value) {
    __threadfence_system();  // the stream-ordered copies before this kernel are complete
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"(value) : "memory");
}

__global__ void k_wait_signals(const uint32_t* flag_a, const uint32_t* flag_b, uint32_t value) {
    const uint32_t* f = threadIdx.x == 0 ? flag_a : flag_b;
    if (f) {
        uint32_t v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v < value) __nanosleep(200);
        } while (v < value);
    }
    __threadfence_system();
}

}  // namespace wc
