// wc_sph_tile.cuh -- warp-cooperative gather kernels for density.comp / update.comp.
//
// One warp owns 32 consecutive cell-sorted particles (one per lane).  Because the array is
// cell-major with x fastest (count.comp:33), the warp's targets that share a (y,z) cell row
// are consecutive lanes, and their joint 27-cell neighbourhood is nine contiguous slices
// [offsets[row'+x0], offsets[row'+x1+1]) of the sorted array.  Per row the warp
//   1. streams each slice once with coalesced float4 loads (32 candidates per iteration),
//   2. culls every candidate against the bounding box of the row's targets grown by h
//      (one lane per candidate, ballot-compacted into a per-warp shared-memory stage),
//   3. runs all 32 targets over the staged survivors with broadcast LDS.128 reads.
// Compared with the thread-per-particle path this turns ~9x32 uncoalesced table walks into
// one coalesced stream, and drops about half of the candidates before the per-pair test.
// The update kernel additionally separates the cheap distance test (phase 1: a bit mask per
// lane) from the expensive pair force (phase 2: lanes walk their own bits), so the heavy
// code runs at near-full lane utilisation instead of once per candidate.
//
// Summation order differs from the simple path (stage order), so floating-point results
// agree to rounding, while neighbour counts and all sort outputs stay bit-exact.
#pragma once

#include "wc_common.cuh"
#include "wc_sph_v1.cuh"

namespace wc {

constexpr int kTileWarps = 8;              // warps (= 32-target groups) per block
constexpr int kChunk = 128;                // staged candidates processed per batch
constexpr int kStageCap = kChunk + 32;     // one cull iteration can overshoot by < 32
constexpr float kFar = 1e18f;              // sentinel coordinate: never within h, no inf/NaN

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct DensityStage {
    float4 a[kStageCap];
};
struct UpdateStage {
    float4 a[kStageCap];
    float4 b[kStageCap];
    uint32_t j[kStageCap];
};

// Per-lane accumulators and the batch processors -----------------------------------------
template <bool kDebug>
struct DensityAcc {
    float sum = 0.0f;
    uint32_t nn = 0;
    // Runs this lane's target over stage[0, count); count is a multiple of 32.
    __device__ __forceinline__ void process(const DensityStage& st, int count, const SphConsts& c,
                                            float4 p, float4, float Teff, uint32_t) {
        for (int k0 = 0; k0 < count; k0 += 32) {
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const float4 q = st.a[k0 + k];
                const float d2 = dist2(p.x - q.x, p.y - q.y, p.z - q.z);
                if (d2 < Teff) {  // density.comp:117; the self pair (d2 = 0) is the m*poly6(0) term
                    sum += poly6_t3(c.h2, d2);
                    if (kDebug) nn++;
                }
            }
        }
    }
};

struct UpdateAcc {
    float Fpx = 0, Fpy = 0, Fpz = 0, Fvx = 0, Fvy = 0, Fvz = 0;
    __device__ __forceinline__ void process(const UpdateStage& st, int count, const SphConsts& c,
                                            float4 p, float4 v, float Teff, uint32_t self) {
        // phase 1: distance test only -> one bit per staged candidate
        unsigned mk[kChunk / 32];
#pragma unroll
        for (int w = 0; w < kChunk / 32; w++) {
            mk[w] = 0u;
            if (w * 32 < count) {
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    const float4 q = st.a[w * 32 + k];
                    const float d2 = dist2(p.x - q.x, p.y - q.y, p.z - q.z);
                    mk[w] |= (d2 < Teff) ? (1u << k) : 0u;
                }
            }
        }
        // phase 2: every lane walks its own accepted candidates (update.comp:174-187)
        unsigned long long m0 = (unsigned long long)mk[0] | ((unsigned long long)mk[1] << 32);
        unsigned long long m1 = (unsigned long long)mk[2] | ((unsigned long long)mk[3] << 32);
        while (__any_sync(0xffffffffu, (m0 | m1) != 0ull)) {
            if ((m0 | m1) != 0ull) {
                int slot;
                if (m0) {
                    slot = __ffsll((long long)m0) - 1;
                    m0 &= m0 - 1ull;
                } else {
                    slot = 64 + __ffsll((long long)m1) - 1;
                    m1 &= m1 - 1ull;
                }
                if (st.j[slot] != self) {  // update.comp:164 (particleID == otherParticleID)
                    const float4 qa = st.a[slot];
                    const float4 qb = st.b[slot];
                    const float rx = p.x - qa.x, ry = p.y - qa.y, rz = p.z - qa.z;
                    pair_force(c, rx, ry, rz, dist2(rx, ry, rz), v.w, v, qa.w, qb, Fpx, Fpy, Fpz,
                               Fvx, Fvy, Fvz);
                }
            }
        }
    }
};
static_assert(kChunk == 128, "UpdateAcc::process packs the masks into two 64-bit words");

// The shared gather driver ----------------------------------------------------------------
// kUpdate selects what is staged (positions only, or positions + velocities + index).
template <bool kUpdate, typename Stage, typename Acc>
__device__ __forceinline__ void gather_rows(const float4* pos_rho, const float4* vel_pres,
                                            const uint32_t* __restrict__ offsets,
                                            const SphConsts& c, Stage& st, Acc& acc, bool valid,
                                            uint32_t self, float4 p, float4 v) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int G = c.G;
    int cx = 0, cy = 0, cz = 0, row = -1;
    if (valid) {
        cx = cell_coord(p.x, c.bin, G), cy = cell_coord(p.y, c.bin, G),
        cz = cell_coord(p.z, c.bin, G);
        row = cz * G + cy;
    }
    const float Tcull = c.T * 1.0001f;  // conservative: rounding in the box distance

    unsigned rem = __ballot_sync(full, valid);
    while (rem) {
        // ---- the lanes of one (y,z) cell row are consecutive (array is cell-sorted)
        const int leader = __ffs(rem) - 1;
        const int r = __shfl_sync(full, row, leader);
        const bool inrow = valid && row == r;
        const unsigned m = __ballot_sync(full, inrow);
        rem &= ~m;
        const int xlo = __shfl_sync(full, cx, __ffs(m) - 1);
        const int xhi = __shfl_sync(full, cx, 31 - __clz(m));
        const int x0 = max(xlo - 1, 0), x1 = min(xhi + 1, G - 1);
        const int ry = __shfl_sync(full, cy, leader), rz = __shfl_sync(full, cz, leader);
        const float bx0 = warp_min_f(inrow ? p.x : INFINITY), bx1 = warp_max_f(inrow ? p.x : -INFINITY);
        const float by0 = warp_min_f(inrow ? p.y : INFINITY), by1 = warp_max_f(inrow ? p.y : -INFINITY);
        const float bz0 = warp_min_f(inrow ? p.z : INFINITY), bz1 = warp_max_f(inrow ? p.z : -INFINITY);
        const float Teff = inrow ? c.T : -1.0f;  // lanes of other rows never accept

        // ---- the nine slices (density.comp:95-101: rows outside the grid are skipped)
        uint32_t sbeg = 0, send = 0;
        if (lane < 9) {
            const int z = rz + lane / 3 - 1, y = ry + lane % 3 - 1;
            if (z >= 0 && z < G && y >= 0 && y < G) {
                const uint32_t rowbase = ((uint32_t)z * G + (uint32_t)y) * G;
                sbeg = offsets[rowbase + x0];
                send = offsets[rowbase + x1 + 1];
            }
        }
        int s = 0, cnt = 0;
        uint32_t j0 = __shfl_sync(full, sbeg, 0), end = __shfl_sync(full, send, 0);
        bool done = false;
        while (true) {
            while (!done && j0 >= end) {
                if (++s == 9) {
                    done = true;
                } else {
                    j0 = __shfl_sync(full, sbeg, s);
                    end = __shfl_sync(full, send, s);
                }
            }
            if (!done) {
                // ---- cull 32 candidates against the targets' box grown by h
                const uint32_t j = j0 + lane;
                const bool ok = j < end;
                float4 q = make_float4(kFar, kFar, kFar, 0.0f);
                if (ok) q = pos_rho[j];
                const float ex = fmaxf(fmaxf(bx0 - q.x, q.x - bx1), 0.0f);
                const float ey = fmaxf(fmaxf(by0 - q.y, q.y - by1), 0.0f);
                const float ez = fmaxf(fmaxf(bz0 - q.z, q.z - bz1), 0.0f);
                const bool keep = ok && (ex * ex + ey * ey + ez * ez < Tcull);
                const unsigned km = __ballot_sync(full, keep);
                if (keep) {
                    const int slot = cnt + __popc(km & lt);
                    st.a[slot] = q;
                    if constexpr (kUpdate) {
                        st.b[slot] = vel_pres[j];
                        st.j[slot] = j;
                    }
                }
                cnt += __popc(km);
                j0 += 32;
            }
            if (cnt >= kChunk || (done && cnt > 0)) {
                int count = kChunk;
                if (cnt < kChunk) {  // final partial batch: pad to a multiple of 32
                    count = (cnt + 31) & ~31;
                    if (cnt + lane < count) st.a[cnt + lane] = make_float4(kFar, kFar, kFar, 0.0f);
                }
                __syncwarp();
                acc.process(st, count, c, p, v, Teff, self);
                __syncwarp();
                // move the (< 32) leftovers to the front
                const int left = cnt - min(cnt, kChunk);
                if (left > 0) {
                    float4 ta, tb;
                    uint32_t tj = 0;
                    const bool mv = lane < left;
                    if (mv) {
                        ta = st.a[kChunk + lane];
                        if constexpr (kUpdate) {
                            tb = st.b[kChunk + lane];
                            tj = st.j[kChunk + lane];
                        }
                    }
                    __syncwarp();
                    if (mv) {
                        st.a[lane] = ta;
                        if constexpr (kUpdate) {
                            st.b[lane] = tb;
                            st.j[lane] = tj;
                        }
                    }
                    __syncwarp();
                }
                cnt = left;
            }
            if (done && cnt == 0) break;
        }
    }
}

template <bool kDebug>
__global__ void __launch_bounds__(kTileWarps * 32)
k_density_tile(float4* pos_rho, float4* __restrict__ vel_pres,
               const uint32_t* __restrict__ offsets, SphConsts c,
               uint32_t* __restrict__ neighbour_counts) {
    __shared__ DensityStage s_stage[kTileWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = (blockIdx.x * kTileWarps + warp) * 32 + lane;
    const bool valid = i < c.n;
    float4 p = make_float4(0, 0, 0, 0);
    if (valid) p = pos_rho[i];
    DensityAcc<kDebug> acc;
    gather_rows<false>(pos_rho, vel_pres, offsets, c, s_stage[warp], acc, valid, (uint32_t)i, p,
                       make_float4(0, 0, 0, 0));
    if (!valid) return;
    float rho, pres;
    finish_density(c, acc.sum, p.x, p.y, p.z, &rho, &pres);
    // In place like density.comp:135; the gather only reads x,y,z, which do not change.
    reinterpret_cast<float*>(pos_rho)[4 * (size_t)i + 3] = rho;
    reinterpret_cast<float*>(vel_pres)[4 * (size_t)i + 3] = pres;
    if (kDebug) neighbour_counts[i] = acc.nn - 1u;  // minus the self pair
}

template <bool kDebug>
__global__ void __launch_bounds__(kTileWarps * 32)
k_update_tile(const float4* __restrict__ pos_rho, const float4* __restrict__ vel_pres,
              const uint32_t* __restrict__ offsets, SphConsts c, float4* __restrict__ pos_out,
              float4* __restrict__ vel_out, float4* __restrict__ forces) {
    __shared__ UpdateStage s_stage[kTileWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = (blockIdx.x * kTileWarps + warp) * 32 + lane;
    const bool valid = i < c.n;
    float4 p = make_float4(0, 0, 0, 0), v = make_float4(0, 0, 0, 0);
    if (valid) {
        p = pos_rho[i];
        v = vel_pres[i];
    }
    UpdateAcc acc;
    gather_rows<true>(pos_rho, vel_pres, offsets, c, s_stage[warp], acc, valid, (uint32_t)i, p, v);
    if (!valid) return;
    const float kp = -(c.m * c.spikyC), kv = c.m * c.viscC;
    float4 po, vo, fo;
    integrate(c, p, v, acc.Fpx * kp, acc.Fpy * kp, acc.Fpz * kp, acc.Fvx * kv, acc.Fvy * kv,
              acc.Fvz * kv, &po, &vo, kDebug ? &fo : nullptr);
    pos_out[i] = po;
    vel_out[i] = vo;
    if (kDebug) forces[i] = fo;
}

inline int tile_blocks(int n) { return (n + kTileWarps * 32 - 1) / (kTileWarps * 32); }

// Returns 0 when launched, -1 when the geometry is not covered (never, currently).
inline int launch_density_tile(float4* pos_rho, float4* vel_pres, const uint32_t* offsets,
                               const SphConsts& c, uint32_t* neighbour_counts,
                               cudaStream_t stream) {
    if (neighbour_counts)
        k_density_tile<true><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, neighbour_counts);
    else
        k_density_tile<false><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, nullptr);
    return 0;
}

inline int launch_update_tile(const float4* pos_rho, const float4* vel_pres,
                              const uint32_t* offsets, const SphConsts& c, float4* pos_out,
                              float4* vel_out, float4* forces, cudaStream_t stream) {
    if (forces)
        k_update_tile<true><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, pos_out, vel_out, forces);
    else
        k_update_tile<false><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, pos_out, vel_out, nullptr);
    return 0;
}

}  // namespace wc
