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

constexpr uint32_t kNoIndex = 0xFFFFFFFFu;     // padding slot in a staged / listed word
constexpr uint32_t kListOverflow = 0xFFFFFFFFu;  // nbr_words value: list did not fit

// self_seq[t] = position (word * 32 + bit) at which target lane t's own particle appears in
// the warp's candidate sequence, recorded when it is staged (density.comp:110: the particle
// itself is not a neighbour, so its bit is cleared from the masks).
struct DensityStage {
    float4 a[kStageCap];
    uint32_t j[kStageCap];
    uint32_t self_seq[32];
};
struct UpdateStage {
    float4 a[kStageCap];
    float4 b[kStageCap];
    uint32_t j[kStageCap];
    uint32_t self_seq[32];
};

// Neighbour list handed from the density pass to the update pass.  For every 32-target
// warp the density kernel records, word by word, which 32 staged candidates it looked at
// (nbr_idx) and which of them each lane accepted (nbr_mask, one bit per candidate).  The
// update kernel replays the words: no second cull, no second distance test.  Layout:
// [(warp * cap_words + word) * 32 + lane], i.e. one coalesced 128-byte line per word.
struct NbrList {
    uint32_t* idx;
    uint32_t* mask;
    uint32_t* words;  // per warp: number of words, or kListOverflow
    int cap_words;
};

// Per-lane accumulators and the batch processors -----------------------------------------
template <bool kDebug>
struct DensityAcc {
    float sum = 0.0f;
    uint32_t nn = 0;
    uint32_t words_used = 0;
    bool overflow = false;
    uint32_t* idx_out = nullptr;   // already offset to this warp's first word + lane
    uint32_t* mask_out = nullptr;
    int cap_words = 0;
    // Runs this lane's target over stage[0, count); count is a multiple of 32.
    __device__ __forceinline__ void process(const DensityStage& st, int count, const SphConsts& c,
                                            float4 p, float4, float Teff, uint32_t) {
        const int lane = threadIdx.x & 31;
        const uint32_t self_seq = st.self_seq[lane];
        for (int k0 = 0; k0 < count; k0 += 32) {
            unsigned mk = 0u;
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const float4 q = st.a[k0 + k];
                const float d2 = dist2(p.x - q.x, p.y - q.y, p.z - q.z);
                if (d2 < Teff) {  // density.comp:117; the self pair (d2 = 0) is the m*poly6(0) term
                    sum += poly6_t3(c.h2, d2);
                    mk |= 1u << k;
                }
            }
            if (words_used == (self_seq >> 5)) mk &= ~(1u << (self_seq & 31u));  // density.comp:110
            if (kDebug) nn += (uint32_t)__popc(mk);
            if (idx_out) {
                if (words_used < (uint32_t)cap_words) {
                    idx_out[(size_t)words_used * 32] = st.j[k0 + lane];
                    mask_out[(size_t)words_used * 32] = mk;
                } else {
                    overflow = true;
                }
            }
            words_used++;
        }
    }
};

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Stage layout of the update pass: a = (x, y, z, 1/rho), b = (vx, vy, vz, P).
struct UpdateAcc {
    // Fp is accumulated without its factor -0.5 * m * spikyC, Fv without m * viscC.
    float Fpx = 0, Fpy = 0, Fpz = 0, Fvx = 0, Fvy = 0, Fvz = 0;
    uint32_t words_used = 0;  // words of the candidate sequence processed so far (full path)

    // One accepted pair of update.comp:174-187.
    __device__ __forceinline__ void pair(const SphConsts& c, float4 p, float4 v, float4 qa,
                                         float4 qb) {
        const float rx = p.x - qa.x, ry = p.y - qa.y, rz = p.z - qa.z;
        const float d2 = dist2(rx, ry, rz);
        const float inv_d = rsqrt_approx(fmaxf(d2, 1e-32f));  // Q7: dist == 0 -> r/d adds 0
        const float hd = c.h - d2 * inv_d;
        const float S = (v.w + qb.w) * qa.w;                   // 2 * (Pi+Pj)/(2 rho_j)
        const float w = S > 0.0f ? (S * (hd * hd)) * inv_d : 0.0f;  // Q9
        Fpx = fmaf(w, rx, Fpx), Fpy = fmaf(w, ry, Fpy), Fpz = fmaf(w, rz, Fpz);
        const float wv = hd * qa.w;                            // update.comp:186-187
        Fvx = fmaf(wv, qb.x - v.x, Fvx), Fvy = fmaf(wv, qb.y - v.y, Fvy),
        Fvz = fmaf(wv, qb.z - v.z, Fvz);
    }

    // phase 2: every lane walks its own accepted candidates.  The four mask words of a batch
    // sit in a per-lane shift register (m0 is the word being drained, base its first slot),
    // so a lane moves on to its next word while others are still busy with theirs.
    __device__ __forceinline__ void pairs(const UpdateStage& st, unsigned m0, unsigned m1,
                                          unsigned m2, unsigned m3, const SphConsts& c, float4 p,
                                          float4 v) {
        int base = 0;
        while (__any_sync(0xffffffffu, (m0 | m1 | m2 | m3) != 0u)) {
            if (m0 == 0u) {
                m0 = m1, m1 = m2, m2 = m3, m3 = 0u;
                base += 32;
            }
            if (m0 != 0u) {
                const int slot = base + __ffs((int)m0) - 1;
                m0 &= m0 - 1u;
                pair(c, p, v, st.a[slot], st.b[slot]);
            }
        }
    }

    // Full path (no list): phase 1 = distance test only -> one bit per staged candidate.
    __device__ __forceinline__ void process(const UpdateStage& st, int count, const SphConsts& c,
                                            float4 p, float4 v, float Teff, uint32_t) {
        const uint32_t self_seq = st.self_seq[threadIdx.x & 31];
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
                if (words_used == (self_seq >> 5)) mk[w] &= ~(1u << (self_seq & 31u));
                words_used++;
            }
        }
        pairs(st, mk[0], mk[1], mk[2], mk[3], c, p, v);
    }
};
static_assert(kChunk == 128, "UpdateAcc::pairs drains four 32-bit mask words per batch");

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
        cz = cell_coord(p.z, c.bin, G) - c.zbase;
        row = cz * G + cy;
    }
    const float Tcull = c.T * 1.0001f;  // conservative: rounding in the box distance
    const uint32_t wfirst = self - (uint32_t)lane;  // index of the warp's first target
    st.self_seq[lane] = kNoIndex;
    __syncwarp();

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
            if (z >= 0 && z < c.Gz && y >= 0 && y < G) {
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
                    if constexpr (kUpdate) {
                        q.w = __frcp_rn(q.w);  // the pair force only needs 1/rho_j
                        st.b[slot] = vel_pres[j];
                    }
                    st.a[slot] = q;
                    st.j[slot] = j;
                    if (j - wfirst < 32u) st.self_seq[j - wfirst] = acc.words_used * 32u + (uint32_t)slot;
                }
                cnt += __popc(km);
                j0 += 32;
            }
            if (cnt >= kChunk || (done && cnt > 0)) {
                int count = kChunk;
                if (cnt < kChunk) {  // final partial batch: pad to a multiple of 32
                    count = (cnt + 31) & ~31;
                    if (cnt + lane < count) {
                        st.a[cnt + lane] = make_float4(kFar, kFar, kFar, 0.0f);
                        st.j[cnt + lane] = kNoIndex;
                    }
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
                        tj = st.j[kChunk + lane];
                        if constexpr (kUpdate) tb = st.b[kChunk + lane];
                    }
                    __syncwarp();
                    if (mv) {
                        st.a[lane] = ta;
                        st.j[lane] = tj;
                        if constexpr (kUpdate) st.b[lane] = tb;
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
               uint32_t* __restrict__ neighbour_counts, NbrList list) {
    __shared__ DensityStage s_stage[kTileWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = blockIdx.x * kTileWarps + warp;
    const int t = wg * 32 + lane;         // target number within this launch
    const int i = c.first + t;            // its index in the candidate arrays
    const bool valid = t < c.n;
    float4 p = make_float4(0, 0, 0, 0);
    if (valid) p = pos_rho[i];
    DensityAcc<kDebug> acc;
    if (list.idx) {
        acc.idx_out = list.idx + (size_t)wg * list.cap_words * 32 + lane;
        acc.mask_out = list.mask + (size_t)wg * list.cap_words * 32 + lane;
        acc.cap_words = list.cap_words;
    }
    gather_rows<false>(pos_rho, vel_pres, offsets, c, s_stage[warp], acc, valid, (uint32_t)i, p,
                       make_float4(0, 0, 0, 0));
    if (list.idx && lane == 0 && wg * 32 < c.n)
        list.words[wg] = acc.overflow ? kListOverflow : acc.words_used;
    if (!valid) return;
    float rho, pres;
    finish_density(c, acc.sum, p.x, p.y, p.z, &rho, &pres);
    // In place like density.comp:135; the gather only reads x,y,z, which do not change.
    reinterpret_cast<float*>(pos_rho)[4 * (size_t)i + 3] = rho;
    reinterpret_cast<float*>(vel_pres)[4 * (size_t)i + 3] = pres;
    if (kDebug) neighbour_counts[t] = acc.nn;  // the self pair's bit is already cleared
}

// update.comp:134-232.  With a valid neighbour list the warp replays the density pass's
// words (gather by index into the stage, then pairs()); without one (list.idx == nullptr,
// or this warp overflowed its list) it runs the full cull + distance test itself.
template <bool kDebug>
__global__ void __launch_bounds__(kTileWarps * 32)
k_update_tile(const float4* __restrict__ pos_rho, const float4* __restrict__ vel_pres,
              const uint32_t* __restrict__ offsets, SphConsts c, float4* __restrict__ pos_out,
              float4* __restrict__ vel_out, float4* __restrict__ forces, NbrList list) {
    __shared__ UpdateStage s_stage[kTileWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = blockIdx.x * kTileWarps + warp;
    const int t = wg * 32 + lane;
    const int i = c.first + t;
    const bool valid = t < c.n;
    float4 p = make_float4(0, 0, 0, 0), v = make_float4(0, 0, 0, 0);
    if (valid) {
        p = pos_rho[i];
        v = vel_pres[i];
    }
    UpdateAcc acc;
    UpdateStage& st = s_stage[warp];
    uint32_t nw = kListOverflow;
    if (list.idx && wg * 32 < c.n) nw = list.words[wg];
    if (nw == kListOverflow) {
        if (wg * 32 < c.n)
            gather_rows<true>(pos_rho, vel_pres, offsets, c, st, acc, valid, (uint32_t)i, p, v);
    } else {
        const uint32_t* widx = list.idx + (size_t)wg * list.cap_words * 32 + lane;
        const uint32_t* wmask = list.mask + (size_t)wg * list.cap_words * 32 + lane;
        for (uint32_t w0 = 0; w0 < nw; w0 += kChunk / 32) {
            unsigned mk[kChunk / 32];
            uint32_t jj[kChunk / 32];
#pragma unroll
            for (int u = 0; u < kChunk / 32; u++) {
                mk[u] = 0u;
                jj[u] = kNoIndex;
                if (w0 + u < nw) {
                    jj[u] = widx[(size_t)(w0 + u) * 32];
                    mk[u] = wmask[(size_t)(w0 + u) * 32];
                }
            }
#pragma unroll
            for (int u = 0; u < kChunk / 32; u++) {
                if (jj[u] != kNoIndex) {
                    float4 qa = pos_rho[jj[u]];
                    qa.w = __frcp_rn(qa.w);
                    st.a[u * 32 + lane] = qa;
                    st.b[u * 32 + lane] = vel_pres[jj[u]];
                }
            }
            __syncwarp();
            acc.pairs(st, mk[0], mk[1], mk[2], mk[3], c, p, v);
            __syncwarp();
        }
    }
    if (!valid) return;
    const float kp = -0.5f * (c.m * c.spikyC), kv = c.m * c.viscC;
    float4 po, vo, fo;
    integrate(c, p, v, acc.Fpx * kp, acc.Fpy * kp, acc.Fpz * kp, acc.Fvx * kv, acc.Fvy * kv,
              acc.Fvz * kv, &po, &vo, kDebug ? &fo : nullptr);
    pos_out[t] = po;
    vel_out[t] = vo;
    if (kDebug) forces[t] = fo;
}

inline int tile_blocks(int n) { return (n + kTileWarps * 32 - 1) / (kTileWarps * 32); }
inline int tile_warps(int n) { return tile_blocks(n) * kTileWarps; }

// Returns 0 when launched, -1 when the geometry is not covered (never, currently).
inline int launch_density_tile(float4* pos_rho, float4* vel_pres, const uint32_t* offsets,
                               const SphConsts& c, uint32_t* neighbour_counts, NbrList list,
                               cudaStream_t stream) {
    if (neighbour_counts)
        k_density_tile<true><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, neighbour_counts, list);
    else
        k_density_tile<false><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, nullptr, list);
    return 0;
}

inline int launch_update_tile(const float4* pos_rho, const float4* vel_pres,
                              const uint32_t* offsets, const SphConsts& c, float4* pos_out,
                              float4* vel_out, float4* forces, NbrList list,
                              cudaStream_t stream) {
    if (forces)
        k_update_tile<true><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, pos_out, vel_out, forces, list);
    else
        k_update_tile<false><<<tile_blocks(c.n), kTileWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, pos_out, vel_out, nullptr, list);
    return 0;
}

}  // namespace wc
