// wc_sph_tile.cuh -- warp-cooperative gather kernels for density.comp / update.comp.
//
// Work unit: a GROUP = up to 32 consecutive cell-sorted particles of ONE (y,z) cell row (one
// per lane; k_build_groups cuts every row into such groups).  Because the array is
// cell-major with x fastest (count.comp:33), the group's joint 27-cell neighbourhood is nine
// contiguous slices [offsets[row'+x0], offsets[row'+x1+1]) of the sorted array.  Per group
// the warp
//   1. streams the nine slices once with coalesced float4 loads (32 candidates per
//      iteration) and culls every candidate against the bounding box of the group's targets
//      grown by h (one lane per candidate, ballot-compacted into a per-warp shared-memory
//      stage in SoA layout),
//   2. runs all 32 targets over the staged survivors: broadcast LDS.128 reads fetch four
//      candidates at a time and the distance test runs on sm_100's packed fp32 pipe
//      (FADD2 / FMUL2 / FFMA2, two candidates per instruction, each half IEEE-rn, so the
//      result is bit-identical to the oracle's fma(rz,rz,fma(ry,ry,rx*rx))),
//   3. records, word by word, which 32 staged candidates it looked at and which of them each
//      lane accepted (the neighbour list), so that
//   4. the update pass replays the list: it re-stages the listed candidates (position,
//      1/rho, velocity, pressure) 256 at a time and every lane walks only its own accepted
//      bits -- no second cull, no second distance test.
// Row-aligned groups keep the culled candidate set near its floor for 32 shared targets
// (~500 at the reference's cell geometry instead of ~720 for unaligned 32-particle runs).
//
// Summation order differs from the simple path, so floating-point results agree to
// rounding, while neighbour counts and all sort outputs stay bit-exact.
#pragma once

#include "wc_common.cuh"
#include "wc_sort.cuh"
#include "wc_sph_v1.cuh"

namespace wc {

#ifndef WC_DENSITY_WARPS
#define WC_DENSITY_WARPS 8
#endif
#ifndef WC_DENSITY_MIN_BLOCKS
#define WC_DENSITY_MIN_BLOCKS 4
#endif
constexpr int kDensityWarps = WC_DENSITY_WARPS;  // warps (= groups) per block, density pass
#ifndef WC_UPDATE_WARPS
#define WC_UPDATE_WARPS 4
#endif
// Tuning knobs (defaults from the sweep recorded in profiles/r01_variant_sweep.md: the update
// pass prefers occupancy over batch size).
#ifndef WC_REPLAY_WORDS
#define WC_REPLAY_WORDS 5
#endif
#ifndef WC_UPDATE_MIN_BLOCKS
#define WC_UPDATE_MIN_BLOCKS 8
#endif
// Pair-loop iterations (two pairs each) per all-lanes-done test of the walk.
#ifndef WC_WALK_UNROLL
#define WC_WALK_UNROLL 1
#endif
// 1: drop every target's own pair from the replayed masks (it adds exactly zero force).
#ifndef WC_DROP_SELF
#define WC_DROP_SELF 1
#endif
// 1: the update pass' list word count goes through a shuffle, so that ptxas can see that it
// is warp-uniform (update_group; without it: convergence checks around the walk, spills).
#ifndef WC_UNIFORM_NW
#define WC_UNIFORM_NW 1
#endif
// 1: the update pass batches the list words strided instead of consecutively (update_group):
// 1-2 % faster, but a lane then sums its pairs in another order than the no-list path does, so
// a group's result depends on whether its list overflowed (list capacity is per handle: the
// bit-identity of z-slabs and whole grid, and of runs with different neighbour_list_words,
// would become conditional).  Off: tests/test_gpu_parity.py::test_neighbour_list_replay_...
#ifndef WC_STRIDED_BATCHES
#define WC_STRIDED_BATCHES 0
#endif
// 1: slab mode, the gathers start in the middle of the group table (see group_prologue).
#ifndef WC_SLAB_MIDDLE_OUT
#define WC_SLAB_MIDDLE_OUT 1
#endif
// 1: list candidates that no target of the group accepted as padding slots.
#ifndef WC_LIST_DROP_UNUSED
#define WC_LIST_DROP_UNUSED 1
#endif
constexpr int kUpdateWarps = WC_UPDATE_WARPS;  // warps per block, update pass (6 KB stage each)
constexpr int kChunk = 128;                // staged candidates per density batch (4 words)
constexpr int kCullDepth = 4;              // density pass: cull loads in flight per lane
constexpr int kRing = 256;                 // density stage: ring of kChunk + 32 * kCullDepth slots
constexpr int kCullOverread = 32 * kCullDepth;  // slack entries behind the sorted arrays (see gather_group)

constexpr int kReplayWords = WC_REPLAY_WORDS;  // list words re-staged per update batch
constexpr int kReplaySlots = kReplayWords * 32;
constexpr float kFar = 1e18f;              // sentinel coordinate: never within h, no inf/NaN

constexpr uint32_t kNoIndex = 0xFFFFFFFFu;       // padding slot in a staged / listed word
constexpr uint32_t kListOverflow = 0xFFFFFFFFu;  // nbr_words value: list did not fit

// ---------------------------------------------------------------------------------------
// Group table.  Every block takes kGroupRows (y,z) rows of the offsets table and cuts each row's
// particle range into groups of <= 32: groups[g] = {index of the group's first particle in
// the sorted arrays, its row (z * G + y of the table), its particle count, 0}.  Blocks claim
// their range of group numbers with one atomicAdd on *num_groups (zero at launch), so the
// numbering is arbitrary between blocks -- nothing depends on it: a group's number only
// selects its slice of the neighbour list, which the density and update passes of the same
// step share.  Rows [row_begin, row_end) are covered (slab mode skips the two ghost layers).
constexpr int kGroupRows = 1024;  // = kBigThreads: the table is built by k_finish_sort's blocks
__device__ __forceinline__ void build_groups_block(const uint32_t* __restrict__ offsets, int G,
                                                   int row_begin, int row_end,
                                                   uint4* __restrict__ groups,
                                                   uint32_t* __restrict__ num_groups) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = row_begin + blockIdx.x * kGroupRows, r = r0 + tid;
    if (tid < 32) s_warp[tid] = 0;
    __syncthreads();
    uint32_t beg = 0, cnt = 0;
    if (r < row_end) {
        beg = offsets[(size_t)r * G];
        cnt = offsets[(size_t)(r + 1) * G] - beg;
    }
    const uint32_t ng = (cnt + 31u) >> 5;
    uint32_t incl = ng;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_base = winc ? atomicAdd(num_groups, winc) : 0u;
    }
    __syncthreads();
    const uint32_t g = s_base + s_warp[warp] + (incl - ng);
    // the warp fills the groups of its 32 rows together (a row can hold many groups)
    for (int src = 0; src < 32; src++) {
        const uint32_t n_src = __shfl_sync(0xffffffffu, cnt, src);
        const uint32_t g_src = __shfl_sync(0xffffffffu, g, src);
        const uint32_t b_src = __shfl_sync(0xffffffffu, beg, src);
        for (uint32_t k = lane; 32u * k < n_src; k += 32)
            groups[g_src + k] = make_uint4(b_src + 32u * k, (uint32_t)(r0 + warp * 32 + src),
                                           min(32u, n_src - 32u * k), 0u);
    }
}

// Last kernel of the sort: the blocks first cut their share of the (y,z) rows into groups,
// then sort and move the cells above kBigCell particles that k_reorder registered (none in a
// physical scene).  One launch for both, because the second part is almost always empty.
// Slab mode with attached neighbours: the reorder (k_reorder and the crowded cells here) has
// stored the halo layers' positions and velocities into the neighbours, so the last block
// raises their "halo positions" flags.
__global__ void __launch_bounds__(kBigThreads)
k_finish_sort(const uint32_t* __restrict__ offsets, int G, int row_begin, int row_end,
              uint4* __restrict__ groups, uint32_t* __restrict__ num_groups,
              uint32_t* __restrict__ ids, uint32_t* __restrict__ scratch, uint32_t base,
              ReorderIO io, SlabRef slab, const uint32_t* __restrict__ big_cells,
              const uint32_t* __restrict__ big_count, uint32_t big_cap, int passes) {
    static_assert(kGroupRows == kBigThreads, "one block shape for both parts");
    bool remote = false;
    if (!slab_dead(slab)) {
        if (groups && row_begin + (int)blockIdx.x * kGroupRows < row_end)
            build_groups_block(offsets, G, row_begin, row_end, groups, num_groups);
        if (ids) {
            const PeerHalo peer = peer_halo_of(slab);
            remote = (peer.pos[0] || peer.pos[1]) && *big_count > 0u;
            reorder_big_cells(ids, scratch, offsets, base, io, peer, big_cells, big_count, big_cap,
                              passes);
        }
    }
    slab_grid_signal(slab, remote, gridDim.x);
}

inline int finish_sort_blocks(int rows) {
    const int for_groups = (rows + kGroupRows - 1) / kGroupRows;
    return for_groups > kBigBlocks ? for_groups : kBigBlocks;
}

// Upper bound of the number of groups for n particles in `rows` rows.
inline int max_groups(int n, long long rows) {
    const long long nonempty = rows < n ? rows : n;
    return (int)((n + 31) / 32 + nonempty);
}

// ---------------------------------------------------------------------------------------
// Packed fp32 pairs (sm_100: add/sub/mul/fma .f32x2, round-to-nearest per half).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Order-preserving float -> uint map, so warp min / max are single REDUX instructions.
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float(u ^ (((int32_t)u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
__device__ __forceinline__ float warp_min_f(float v, bool use) {
    return ord2f(__reduce_min_sync(0xffffffffu, use ? f2ord(v) : 0xFFFFFFFFu));
}
__device__ __forceinline__ float warp_max_f(float v, bool use) {
    return ord2f(__reduce_max_sync(0xffffffffu, use ? f2ord(v) : 0u));
}

// Index of the most significant set bit (0xffffffff for 0), and the mask of the bits below an
// index (all ones for an index >= 32): one instruction each (FLO, BMSK).
__device__ __forceinline__ unsigned highest_bit(unsigned m) {
    unsigned r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(m));
    return r;
}
__device__ __forceinline__ unsigned bits_below(unsigned idx) {
    unsigned r;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0u), "r"(idx));
    return r;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------
// Stages.  A target's own particle is one of the staged candidates: it passes the distance
// test with d2 = 0, which is exactly the m * poly6(0) term the density starts from
// (density.comp:92), and its pair force is exactly zero (r = 0, v_j - v_i = 0), so it stays in
// the accept masks; only the neighbour COUNT subtracts it (density.comp:110, j != i).
struct alignas(16) DensityStage {
    static constexpr int kBatch = kChunk;
    static constexpr int kDepth = kCullDepth;
    static constexpr int kWrap = kRing - 1;  // ring: batches start at slot 0 or kChunk
    float x[kRing];
    float y[kRing];
    float z[kRing];
    uint32_t j[kRing];
    __device__ __forceinline__ void put(int slot, float4 q, uint32_t j_, const float4*) {
        x[slot] = q.x, y[slot] = q.y, z[slot] = q.z, j[slot] = j_;
    }
    __device__ __forceinline__ void init(int) {}
    __device__ __forceinline__ void pad(int slot) {
        x[slot] = kFar, y[slot] = kFar, z[slot] = kFar, j[slot] = kNoIndex;
    }
};
static_assert(kRing == 2 * kChunk && kChunk + 32 * kCullDepth <= kRing, "ring sizing");

// Update pass: a = (x, y, z, 1/rho), b = (vx, vy, vz, P); mask[w * 32 + lane] = the bits of
// word w accepted by `lane`, row kReplayWords is a zero terminator (init() writes it once
// per warp).  Sized for a replay batch; the no-list path uses the first kBatch + 32 slots.
struct alignas(16) UpdateStage {
    // no-list path: linear stage of kBatch + 32 slots inside the replay buffers, depth 1
    static constexpr int kBatch = (kReplayWords - 1) * 32 < kChunk ? (kReplayWords - 1) * 32 : kChunk;
    static constexpr int kDepth = 1;
    static constexpr int kWrap = 0;  // linear stage: leftovers are moved to the front
    float4 a[kReplaySlots];
    float4 b[kReplaySlots];  // walk() relies on b directly following a
    uint32_t mask[(kReplayWords + 1) * 32];
    __device__ __forceinline__ void init(int lane) { mask[kReplayWords * 32 + lane] = 0u; }
    __device__ __forceinline__ void put(int slot, float4 q, uint32_t j_, const float4* vel_pres) {
        q.w = __frcp_rn(q.w);  // the pair force only needs 1/rho_j
        a[slot] = q;
        b[slot] = vel_pres[j_];
    }
    __device__ __forceinline__ void pad(int slot) { a[slot] = make_float4(kFar, kFar, kFar, 0.0f); }
    __device__ __forceinline__ void move(int dst, int src, bool mv) {
        float4 ta, tb;
        if (mv) ta = a[src], tb = b[src];
        __syncwarp();
        if (mv) a[dst] = ta, b[dst] = tb;
    }
};
static_assert(UpdateStage::kBatch >= 32 && UpdateStage::kBatch + 32 <= kReplaySlots,
              "the no-list path stages into the replay buffers");

// Neighbour list handed from the density pass to the update pass.  Layout:
// [(group * cap_words + word) * 32 + lane], i.e. one coalesced 128-byte line per word, for
// both the candidate indices (idx) and the per-lane accept masks (mask).
struct NbrList {
    uint32_t* idx;
    uint32_t* mask;
    uint32_t* words;  // per group: number of words, or kListOverflow
    int cap_words;
};

// ---------------------------------------------------------------------------------------
// Density accumulator: phase 2 of the file comment.
template <bool kDebug>
struct DensityAcc {
    float sum0 = 0.0f, sum1 = 0.0f;
    uint32_t nn = 0;
    uint32_t words_used = 0;
    bool overflow = false;
    uint32_t* idx_out = nullptr;   // already offset to this group's first word + lane
    uint32_t* mask_out = nullptr;
    int cap_words = 0;

    // Runs this lane's target over stage[head, head + count); count is a multiple of 32.
    __device__ __forceinline__ void process(const DensityStage& st, int head, int count,
                                            const SphConsts& c, float4 p, float4, float Teff) {
        const int lane = threadIdx.x & 31;
        const f32x2 PX = pack2(p.x, p.x), PY = pack2(p.y, p.y), PZ = pack2(p.z, p.z);
        const f32x2 H2 = pack2(c.h2, c.h2);
        for (int k0 = head; k0 < head + count; k0 += 32) {
            unsigned mk = 0u;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 X = *reinterpret_cast<const float4*>(&st.x[k0 + 4 * q]);
                const float4 Y = *reinterpret_cast<const float4*>(&st.y[k0 + 4 * q]);
                const float4 Z = *reinterpret_cast<const float4*>(&st.z[k0 + 4 * q]);
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const f32x2 rx = sub2(PX, hh ? pack2(X.z, X.w) : pack2(X.x, X.y));
                    const f32x2 ry = sub2(PY, hh ? pack2(Y.z, Y.w) : pack2(Y.x, Y.y));
                    const f32x2 rz = sub2(PZ, hh ? pack2(Z.z, Z.w) : pack2(Z.x, Z.y));
                    // fma(rz,rz, fma(ry,ry, rx*rx)) per half: the oracle's pinned op order
                    const f32x2 d2 = fma2(rz, rz, fma2(ry, ry, mul2(rx, rx)));
                    const f32x2 t = sub2(H2, d2);       // density.comp:54, on squared distances
                    const f32x2 tt = mul2(t, t);
                    float d2a, d2b, ta, tb, tta, ttb;
                    unpack2(d2, d2a, d2b);
                    unpack2(t, ta, tb);
                    unpack2(tt, tta, ttb);
                    // density.comp:117; the self pair (d2 = 0) is the m * poly6(0) term
                    if (d2a < Teff) {
                        sum0 = fmaf(tta, ta, sum0);
                        mk |= 1u << (4 * q + 2 * hh);
                    }
                    if (d2b < Teff) {
                        sum1 = fmaf(ttb, tb, sum1);
                        mk |= 1u << (4 * q + 2 * hh + 1);
                    }
                }
            }
            if (kDebug) nn += (uint32_t)__popc(mk);
            if (idx_out) {
                if (words_used < (uint32_t)cap_words) {
#if WC_LIST_DROP_UNUSED
                    // a staged candidate no target accepted (a third of them: the cull keeps the
                    // box of the targets grown by h) is listed as a padding slot, so the update
                    // pass does not fetch it
                    const unsigned any = __reduce_or_sync(0xffffffffu, mk);
                    idx_out[(size_t)words_used * 32] = ((any >> lane) & 1u) ? st.j[k0 + lane] : kNoIndex;
#else
                    idx_out[(size_t)words_used * 32] = st.j[k0 + lane];
#endif
                    mask_out[(size_t)words_used * 32] = mk;
                } else {
                    overflow = true;
                }
            }
            words_used++;
        }
    }
};

// Update accumulator.  kExt: also the colour-field sums of the surface tension (extended
// physics; their own instantiation, so the reference step's walk loop carries none of it).
template <bool kExt>
struct UpdateAcc {
    ColourField cf;
    // Fp is accumulated without its factor -0.5 * m * spikyC, Fv without m * viscC.
    float Fpx = 0, Fpy = 0, Fpz = 0, Fvx = 0, Fvy = 0, Fvz = 0;
    uint32_t words_used = 0;  // words of the candidate sequence processed so far (no-list path)

    // One accepted pair of update.comp:174-187.  A lane without a pick passes stale operands
    // with 1/rho_j forced to zero, which makes the pair a no-op (`on` is informative only).
    __device__ __forceinline__ void pair(const ConstsOf<kExt>& c, float4 p, float4 v, float4 qa,
                                         float4 qb, bool on) {
        // x and y of the separation and of the velocity difference as packed pairs: the operands
        // already sit in aligned register pairs (LDS.128 / LDG.128 results)
        float rx, ry, dvx, dvy;
        unpack2(sub2(pack2(p.x, p.y), pack2(qa.x, qa.y)), rx, ry);
        unpack2(sub2(pack2(qb.x, qb.y), pack2(v.x, v.y)), dvx, dvy);
        const float rz = p.z - qa.z;
        const float d2 = dist2(rx, ry, rz);
        const float inv_d = rsqrt_approx(fmaxf(d2, 1e-32f));  // Q7: dist == 0 -> r/d adds 0
        const float hd = c.h - d2 * inv_d;
        // A lane without a pick has 1/rho_j = 0 (pick() zeroes it): S = 0 fails the test below
        // and wv = 0, so the stale pair adds exactly zero without any select on `on`.
        const float S = (v.w + qb.w) * qa.w;                   // 2 * (Pi+Pj)/(2 rho_j)
        const float w = (S > 0.0f) ? (S * (hd * hd)) * inv_d : 0.0f;  // Q9
        Fpx = fmaf(w, rx, Fpx), Fpy = fmaf(w, ry, Fpy), Fpz = fmaf(w, rz, Fpz);
        const float wv = hd * qa.w;                            // update.comp:186-187
        Fvx = fmaf(wv, dvx, Fvx), Fvy = fmaf(wv, dvy, Fvy), Fvz = fmaf(wv, qb.z - v.z, Fvz);
        if constexpr (kExt) cf.add(c, rx, ry, rz, d2, qa.w);  // (1/rho_j = 0 without a pick)
    }

    // Every lane walks its own accepted bits of the staged mask words [0, nwb) -- one flat
    // loop over the whole batch, two pairs per iteration, so lanes only wait for each other
    // at the batch's end.  The lane keeps the word it is draining (m) and the next one (mn,
    // prefetched) in registers, so moving on to the next word is a few predicated
    // instructions and its shared-memory load is off the critical path; a pick that finds
    // no bit (empty word, or lane finished) loads nothing and adds zero.
    // Mask rows [nwb, kReplayWords] must be zero.
    __device__ __forceinline__ void walk(const UpdateStage& st, int nwb, const ConstsOf<kExt>& c,
                                         float4 p, float4 v) {
        const int lane = threadIdx.x & 31;
        // the mask words are addressed through ONE 32-bit shared-memory address (a generic pointer
        // costs a second induction variable for the end test)
        uint32_t mptr = (uint32_t)__cvta_generic_to_shared(st.mask + lane);
        const uint32_t mend = mptr + 128u * (uint32_t)nwb;  // a zero row
        auto lds_u32 = [](uint32_t a) -> unsigned {
            unsigned r;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
            return r;
        };
        unsigned m = lds_u32(mptr), mn = (nwb > 1) ? lds_u32(mptr + 128u) : 0u;
        mptr = (nwb > 2) ? mptr + 256u : mend;
        const float4* abase = st.a;  // slot 0 of the word being drained
        float4 qa0 = make_float4(0, 0, 0, 0), qb0 = qa0, qa1 = qa0, qb1 = qa0;
        auto pick = [&](float4& qa, float4& qb) -> bool {
            if (m == 0u) {
                m = mn;
                abase += 32;
                mn = lds_u32(mptr);
                mptr = (mptr == mend) ? mend : mptr + 128u;
            }
            const bool on = m != 0u;
            const unsigned hb = highest_bit(m);  // 0xffffffff when no bit is left
            if (on) {
                const float4* q = abase + hb;
                qa = q[0];
                qb = q[kReplaySlots];
            } else {
                qa.w = 0.0f;  // see pair()
            }
            m &= bits_below(hb);  // drops bit hb (m stays 0 when it was 0)
            return on;
        };
        while (__any_sync(0xffffffffu, (m | mn) != 0u || mptr != mend)) {
#pragma unroll
            for (int rep = 0; rep < WC_WALK_UNROLL; rep++) {
                const bool on0 = pick(qa0, qb0);
                const bool on1 = pick(qa1, qb1);
                pair(c, p, v, qa0, qb0, on0);
                pair(c, p, v, qa1, qb1, on1);
            }
        }
    }

    // No-list path: phase 1 = distance test only -> one bit per staged candidate, then walk.
    __device__ __forceinline__ void process(UpdateStage& st, int /*head = 0*/, int count,
                                            const ConstsOf<kExt>& c, float4 p, float4 v,
                                            float Teff) {
        const int lane = threadIdx.x & 31;
        const int nw = count >> 5;
        for (int w = 0; w < nw; w++) {
            unsigned mk = 0u;
#pragma unroll 8
            for (int k = 0; k < 32; k++) {
                const float4 q = st.a[w * 32 + k];
                const float d2 = dist2(p.x - q.x, p.y - q.y, p.z - q.z);
                mk |= (d2 < Teff) ? (1u << k) : 0u;
            }
            words_used++;
            st.mask[w * 32 + lane] = mk;
        }
        for (int w = nw; w < kReplayWords; w++) st.mask[w * 32 + lane] = 0u;
        __syncwarp();
        walk(st, nw, c, p, v);
    }
};

// ---------------------------------------------------------------------------------------
// The shared gather driver: phase 1 (cull + stage) and the hand-over to acc.process().
struct GroupGeom {
    int x0, x1;        // cell columns of the nine slices
    int ry, rz;        // the group's cell row (rz relative to the table's layer 0)
};

template <typename Stage, typename Acc, typename Consts>
__device__ __forceinline__ void gather_group(const float4* pos_rho, const float4* vel_pres,
                                             const uint32_t* __restrict__ offsets,
                                             const Consts& c, Stage& st, Acc& acc, bool valid,
                                             const GroupGeom& gg, float4 p,
                                             float4 v) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int G = c.G;
    // NaN targets never accept anything; keep them out of the box so it stays finite.
    const bool box = valid && p.x == p.x && p.y == p.y && p.z == p.z;
    const float bx0 = warp_min_f(p.x, box), bx1 = warp_max_f(p.x, box);
    const float by0 = warp_min_f(p.y, box), by1 = warp_max_f(p.y, box);
    const float bz0 = warp_min_f(p.z, box), bz1 = warp_max_f(p.z, box);
    const float Tcull = c.T * 1.0001f;  // conservative: rounding in the box distance
    const float Teff = valid ? c.T : -1.0f;
    st.init(lane);
    __syncwarp();

    // the nine slices (density.comp:95-101: rows outside the grid are skipped)
    uint32_t sbeg = 0, send = 0;
    if (lane < 9) {
        const int z = gg.rz + lane / 3 - 1, y = gg.ry + lane % 3 - 1;
        if (z >= 0 && z < c.Gz && y >= 0 && y < G) {
            const uint32_t rowbase = ((uint32_t)z * G + (uint32_t)y) * G;
            sbeg = offsets[rowbase + gg.x0];
            send = offsets[rowbase + gg.x1 + 1];
        }
    }
    constexpr int kDepth = Stage::kDepth;
    int cnt = 0, head = 0;  // pending candidates are stage[head, head + cnt) (mod ring)
    for (int s = 0; s < 9; s++) {
        const uint32_t end = __shfl_sync(full, send, s);
        for (uint32_t j0 = __shfl_sync(full, sbeg, s); j0 < end; j0 += 32 * kDepth) {
            // kDepth x 32 candidates in flight, then culled against the targets' box grown
            // by h, 32 at a time, and ballot-compacted behind the pending ones
            // The loads are unconditional (one base address, immediate offsets): a slice's last
            // iteration reads up to kCullOverread - 1 entries past its end, which the buffers
            // allow for; those lanes are dropped by the index test below.
            float4 q[kDepth];
            const float4* src = pos_rho + j0 + lane;
#pragma unroll
            for (int k = 0; k < kDepth; k++) q[k] = src[32 * k];
#pragma unroll
            for (int k = 0; k < kDepth; k++) {
                const uint32_t j = j0 + 32 * k + lane;
                const float ex = fmaxf(fmaxf(bx0 - q[k].x, q[k].x - bx1), 0.0f);
                const float ey = fmaxf(fmaxf(by0 - q[k].y, q[k].y - by1), 0.0f);
                const float ez = fmaxf(fmaxf(bz0 - q[k].z, q[k].z - bz1), 0.0f);
                const bool keep = j < end && ex * ex + ey * ey + ez * ez < Tcull;
                const unsigned km = __ballot_sync(full, keep);
                if (keep) {
                    const int at = cnt + __popc(km & lt);
                    st.put(Stage::kWrap ? ((head + at) & Stage::kWrap) : at, q[k], j, vel_pres);
                }
                cnt += __popc(km);
            }
            if (cnt >= Stage::kBatch) {
                __syncwarp();
                acc.process(st, head, Stage::kBatch, c, p, v, Teff);
                __syncwarp();
                cnt -= Stage::kBatch;
                if constexpr (Stage::kWrap != 0) {
                    head ^= Stage::kBatch;  // the ring's other half
                } else if (cnt > 0) {  // linear stage: move the (< 32) leftovers to the front
                    st.move(lane, Stage::kBatch + lane, lane < cnt);
                    __syncwarp();
                }
            }
        }
    }
    if (cnt > 0) {  // final partial batch: pad to a multiple of 32
        const int count = (cnt + 31) & ~31;
        if (cnt + lane < count) st.pad(head + cnt + lane);
        __syncwarp();
        acc.process(st, head, count, c, p, v, Teff);
        __syncwarp();
    }
}

// Hash + count of the NEXT step's sort, done by the update pass on the positions it has just
// integrated (count.comp:25-36 on this step's output): the next step then starts at the scan.
// counts: the next step's (zeroed) cell counters; cell_ids / ranks: per output particle, what
// k_hash_count would have written.  All null: not wanted.
struct PreHash {
    uint32_t* counts;
    uint32_t* cell_ids;
    uint32_t* ranks;
};

// Common prologue: which group this warp owns, its targets and row geometry.
struct GroupCtx {
    int g;           // group number
    int t;           // this lane's target number within the launch (output index)
    int i;           // this lane's particle index in the candidate arrays
    bool active;     // the warp has a group
    bool valid;      // this lane has a target
    int cnt;         // targets of the group (lanes 0 .. cnt - 1 are valid)
    GroupGeom gg;
};

template <bool kSlab>
__device__ __forceinline__ GroupCtx group_prologue(const float4* pos_rho, const SphConsts& c,
                                                   const uint4* __restrict__ groups,
                                                   const uint32_t* __restrict__ num_groups,
                                                   int warps_per_block, float4* p_out,
                                                   unsigned vblock, unsigned needed_blocks) {
    GroupCtx x;
    const int lane = threadIdx.x & 31;
    // Slab mode: the (virtual) blocks start in the middle of the group table and wrap around,
    // so the groups of the first / last owned layer -- the only ones that wait for a
    // neighbour's halo -- come up about half-way through the kernel, when the halo has long
    // arrived.
    unsigned block = vblock;
    if constexpr (kSlab && WC_SLAB_MIDDLE_OUT) {
        block = block + needed_blocks / 2u;
        block = block >= needed_blocks ? block - needed_blocks : block;
    }
    x.g = (int)block * warps_per_block + (threadIdx.x >> 5);
    const uint4 rec = groups[x.g];  // the table is allocated for the launch bound; may be stale
    x.active = (uint32_t)x.g < *num_groups;
    // (slab mode: an idle warp of the last block does not exit, it goes on to the block's
    // signal; the vote lets ptxas see that the branch around the gather is warp-uniform, see
    // slab_warp_wait); a block beyond the table must not wrap around into it
    if constexpr (kSlab) x.active = __any_sync(0xffffffffu, x.active && vblock < needed_blocks);
    x.valid = false;
    x.t = x.i = x.cnt = 0;
    *p_out = make_float4(0, 0, 0, 0);
    if (!x.active) return x;
    const int G = c.G;
    x.gg.rz = (int)(rec.y / (uint32_t)G);
    x.gg.ry = (int)(rec.y - (uint32_t)x.gg.rz * (uint32_t)G);
    const int cnt = (int)rec.z;  // >= 1
    x.valid = lane < cnt;
    x.cnt = cnt;
    x.i = (int)rec.x + lane;
    x.t = x.i - c.first;
    float4 p = make_float4(0, 0, 0, 0);
    if (x.valid) p = pos_rho[x.i];
    *p_out = p;
    // cells ascend along the row, so the first / last valid lanes bound the x range
    const int cx = x.valid ? cell_coord(p.x, c.bin, G) : 0;
    const int xlo = __shfl_sync(0xffffffffu, cx, 0), xhi = __shfl_sync(0xffffffffu, cx, cnt - 1);
    x.gg.x0 = max(xlo - 1, 0);
    x.gg.x1 = min(xhi + 1, G - 1);
    return x;
}

template <bool kDebug, bool kSlab, bool kExt>
__device__ __forceinline__ bool density_group(float4* pos_rho, float4* __restrict__ vel_pres,
                                              const uint32_t* __restrict__ offsets,
                                              const ConstsOf<kExt>& c,
                                              const uint4* __restrict__ groups,
                                              const uint32_t* __restrict__ num_groups,
                                              uint32_t* __restrict__ neighbour_counts,
                                              const NbrList& list, const SlabRef& slab,
                                              DensityStage& stage, unsigned vblock,
                                              unsigned needed_blocks) {
    const int lane = threadIdx.x & 31;
    float4 p;
    const GroupCtx x = group_prologue<kSlab>(pos_rho, c, groups, num_groups, kDensityWarps, &p,
                                             vblock, needed_blocks);
    if (!x.active) return false;
    // slab mode: only the groups of the first / last owned layer read the neighbours' halo
    // positions (stored into this rank's ghost slots by their reorder)
    if constexpr (kSlab) slab_warp_wait(slab, x.gg.rz <= 1, x.gg.rz >= c.Gz - 2);
    DensityAcc<kDebug> acc;
    if (list.idx) {
        acc.idx_out = list.idx + (size_t)x.g * list.cap_words * 32 + lane;
        acc.mask_out = list.mask + (size_t)x.g * list.cap_words * 32 + lane;
        acc.cap_words = list.cap_words;
    }
    gather_group(pos_rho, vel_pres, offsets, c, stage, acc, x.valid, x.gg, p,
                 make_float4(0, 0, 0, 0));
    if (list.idx && lane == 0) list.words[x.g] = acc.overflow ? kListOverflow : acc.words_used;
    if (!x.valid) return false;
    float rho, pres;
    finish_density<kExt>(c, acc.sum0 + acc.sum1, p.x, p.y, p.z, &rho, &pres);
    bool remote = false;
    // In place like density.comp:135; the gather only reads x,y,z, which do not change.
    reinterpret_cast<float*>(pos_rho)[4 * (size_t)x.i + 3] = rho;
    reinterpret_cast<float*>(vel_pres)[4 * (size_t)x.i + 3] = pres;
    // a halo particle: density and pressure also go into the neighbour's ghost copy (whose
    // x,y,z and velocity the reorder already stored there)
    if constexpr (kSlab) {
        const PeerHalo peer = peer_halo_of(slab);
        const uint32_t t = (uint32_t)x.t;
        if (peer.pos[0] && t < peer.n_first) {
            reinterpret_cast<float*>(peer.pos[0])[4 * (size_t)(peer.dst[0] + t) + 3] = rho;
            reinterpret_cast<float*>(peer.vel[0])[4 * (size_t)(peer.dst[0] + t) + 3] = pres;
            remote = true;
        }
        if (peer.pos[1] && t >= peer.hi_begin) {
            reinterpret_cast<float*>(peer.pos[1])[4 * (size_t)(peer.dst[1] + t - peer.hi_begin) + 3] = rho;
            reinterpret_cast<float*>(peer.vel[1])[4 * (size_t)(peer.dst[1] + t - peer.hi_begin) + 3] = pres;
            remote = true;
        }
    }
    if (kDebug) {  // the self pair was accepted iff the particle's own d2 is 0 (finite position)
        const bool self = dist2(p.x - p.x, p.y - p.y, p.z - p.z) < c.T;
        neighbour_counts[x.t] = acc.nn - (self ? 1u : 0u);
    }
    return remote;
}

// density.comp:81-137, one warp per group.  Slab mode: the launch is sized by capacity, so
// the blocks beyond the group table leave at once; with attached neighbours the warps of the
// boundary layers wait for the halo positions, and once every block is done the neighbours
// are told that this rank's halo density / pressure is in their ghost copies.
template <bool kDebug, bool kSlab, bool kExt>
__global__ void __launch_bounds__(kDensityWarps * 32, WC_DENSITY_MIN_BLOCKS)
k_density_tile(float4* pos_rho, float4* __restrict__ vel_pres,
               const uint32_t* __restrict__ offsets, ConstsOf<kExt> c,
               const uint4* __restrict__ groups, const uint32_t* __restrict__ num_groups,
               uint32_t* __restrict__ neighbour_counts, NbrList list, SlabRef slab) {
    __shared__ DensityStage s_stage[kDensityWarps];
    if constexpr (!kSlab) {  // whole grid: no slab code at all in this instantiation
        density_group<kDebug, false, kExt>(pos_rho, vel_pres, offsets, c, groups, num_groups,
                                     neighbour_counts, list, slab, s_stage[threadIdx.x >> 5],
                                     blockIdx.x, gridDim.x);
    } else {
        // at least one block stays to raise the signal, also in a dead step
        const uint32_t blocks = max(1u, (*num_groups + kDensityWarps - 1u) / kDensityWarps);
        if (blockIdx.x >= blocks) return;
        bool remote = false;
        if (!slab_dead(slab))
            remote = density_group<kDebug, true, kExt>(pos_rho, vel_pres, offsets, c, groups, num_groups,
                                                 neighbour_counts, list, slab,
                                                 s_stage[threadIdx.x >> 5], blockIdx.x, blocks);
        slab_grid_signal(slab, remote, blocks);
    }
}

// update.comp:134-232.  With a valid neighbour list the warp replays the density pass's
// words; without one (list.idx == nullptr, or this group overflowed its list) it runs the
// cull + distance test itself.
template <bool kDebug, bool kSlab, bool kExt>
__device__ __forceinline__ void update_group(const float4* __restrict__ pos_rho,
                                             const float4* __restrict__ vel_pres,
                                             const uint32_t* __restrict__ offsets,
                                             const ConstsOf<kExt>& c,
                                             const uint4* __restrict__ groups,
                                             const uint32_t* __restrict__ num_groups,
                                             float4* __restrict__ pos_out,
                                             float4* __restrict__ vel_out,
                                             float4* __restrict__ forces, const NbrList& list,
                                             float4* __restrict__ aos_out, const SlabRef& slab,
                                             UpdateStage& st, unsigned vblock,
                                             unsigned needed_blocks, const PreHash& pre) {
    const int lane = threadIdx.x & 31;
    float4 p;
    const GroupCtx x = group_prologue<kSlab>(pos_rho, c, groups, num_groups, kUpdateWarps, &p,
                                             vblock, needed_blocks);
    if (!x.active) return;
    // slab mode with attached neighbours: the ghosts' density / pressure must have arrived
    // before a group of the first / last owned layer reads them
    if constexpr (kSlab) slab_warp_wait(slab, x.gg.rz <= 1, x.gg.rz >= c.Gz - 2);
    float4 v = make_float4(0, 0, 0, 0);
    if (x.valid) v = vel_pres[x.i];
    UpdateAcc<kExt> acc;
    uint32_t nw = kListOverflow;
    if (list.idx) nw = list.words[x.g];
#if WC_UNIFORM_NW
    nw = __shfl_sync(0xffffffffu, nw, 0);  // the same word in every lane; this way ptxas knows it too
#endif
    if (nw == kListOverflow) {
        gather_group(pos_rho, vel_pres, offsets, c, st, acc, x.valid, x.gg, p, v);
    } else {
        const uint32_t* widx = list.idx + (size_t)x.g * list.cap_words * 32 + lane;
        const uint32_t* wmask = list.mask + (size_t)x.g * list.cap_words * 32 + lane;
        const uint32_t first_target = (uint32_t)(x.i - lane);
        st.init(lane);
#if WC_STRIDED_BATCHES
        // The nw list words go into nb = ceil(nw / K) batches, batch b taking the words b, b + nb,
        // b + 2 nb, ...: equal-sized batches, each a cross-section of the nine slices.  The walk
        // ends a batch with the lane that accepted most of it, and consecutive words (one
        // slice: one side of the group) are accepted unevenly -- tests/model/model_walk.py:
        // 37.7 instead of 41.0 pair-loop iterations per group.
        const uint32_t nb = (nw + (uint32_t)kReplayWords - 1u) / (uint32_t)kReplayWords;
        for (uint32_t b = 0; b < nb; b++) {
            uint32_t jj[kReplayWords];
            int nwb = 0;
#pragma unroll
            for (int u = 0; u < kReplayWords; u++) {
                const uint32_t w = b + (uint32_t)u * nb;
                jj[u] = kNoIndex;
                uint32_t mk = 0u;
                if (w < nw) {
                    jj[u] = widx[(size_t)w * 32];
                    mk = wmask[(size_t)w * 32];
                    nwb = u + 1;
                }
                st.mask[u * 32 + lane] = mk;
            }
#else
        for (uint32_t w0 = 0; w0 < nw; w0 += kReplayWords) {
            uint32_t jj[kReplayWords];
            const int nwb = (int)min((uint32_t)kReplayWords, nw - w0);
#pragma unroll
            for (int u = 0; u < kReplayWords; u++) {
                jj[u] = kNoIndex;
                uint32_t mk = 0u;
                if (w0 + u < nw) {
                    jj[u] = widx[(size_t)(w0 + u) * 32];
                    mk = wmask[(size_t)(w0 + u) * 32];
                }
                st.mask[u * 32 + lane] = mk;
            }
#endif
#pragma unroll
            for (int u = 0; u < kReplayWords; u++) {
                if (jj[u] != kNoIndex) {
                    float4 qa = pos_rho[jj[u]];
                    qa.w = __frcp_rn(qa.w);
                    st.a[u * 32 + lane] = qa;
                    st.b[u * 32 + lane] = vel_pres[jj[u]];
                }
            }
            __syncwarp();
#if WC_DROP_SELF
            // a listed candidate that is one of this group's own targets: drop that target's
            // self pair (it would add exactly zero; this only saves the evaluation)
#pragma unroll
            for (int u = 0; u < kReplayWords; u++) {
                const uint32_t t = jj[u] - first_target;
                if (t < 32u) atomicAnd(&st.mask[u * 32 + t], ~(1u << lane));
            }
            __syncwarp();
#endif
            acc.walk(st, nwb, c, p, v);
            __syncwarp();
        }
    }
    if (!x.valid) return;
    const float kp = -0.5f * (c.m * c.spikyC), kv = c.m * c.viscC;
    float4 po, vo, fo;
    integrate<kExt>(c, p, v, acc.Fpx * kp, acc.Fpy * kp, acc.Fpz * kp, acc.Fvx * kv, acc.Fvy * kv,
                    acc.Fvz * kv, &po, &vo, kDebug ? &fo : nullptr, acc.cf);
    pos_out[x.t] = po;
    vel_out[x.t] = vo;
    if (kDebug) forces[x.t] = fo;
    // wc_step_host: the same record also goes out as the reference's 32-byte AoS Particle
    // (util.h:29-35), straight into the caller's mapped pinned host buffer -- the
    // device-to-host copy rides on the kernel instead of following it (a warp writes 1 KB
    // contiguous).
    if (aos_out) {
        __stcs(&aos_out[2 * (size_t)x.t], po);
        __stcs(&aos_out[2 * (size_t)x.t + 1], vo);
    }
    if (pre.counts) {  // k_hash_count for the next step, on the position just stored
        const unsigned mine = x.cnt >= 32 ? 0xffffffffu : ((1u << x.cnt) - 1u);
        uint32_t cn;
        if constexpr (kSlab) {
            // a particle that left the slab's layers was packed for the neighbour (k_migrants):
            // it takes no part in this slab's next sort (k_hash_count_slab)
            const int cz = cell_coord(po.z, c.bin, c.G) - c.zbase;
            cn = (cz >= 1 && cz <= c.Gz - 2) ? cell_index(po.x, po.y, po.z, c.bin, c.G, c.zbase)
                                              : 0xFFFFFFFFu;
        } else {
            cn = cell_index(po.x, po.y, po.z, c.bin, c.G);
        }
        const unsigned same = __match_any_sync(mine, cn);
        pre.cell_ids[x.t] = cn;
        if (cn != 0xFFFFFFFFu) {
            const int leader = __ffs(same) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&pre.counts[cn], (uint32_t)__popc(same));
            base = __shfl_sync(same, base, leader);
            pre.ranks[x.t] = base + (uint32_t)__popc(same & ((1u << lane) - 1u));
        }
    }
}

// (slab mode: launch sizing as in k_density_tile)
template <bool kDebug, bool kSlab, bool kExt>
__global__ void __launch_bounds__(kUpdateWarps * 32, WC_UPDATE_MIN_BLOCKS)
k_update_tile(const float4* __restrict__ pos_rho, const float4* __restrict__ vel_pres,
              const uint32_t* __restrict__ offsets, ConstsOf<kExt> c,
              const uint4* __restrict__ groups, const uint32_t* __restrict__ num_groups,
              float4* __restrict__ pos_out, float4* __restrict__ vel_out,
              float4* __restrict__ forces, NbrList list, float4* __restrict__ aos_out,
              SlabRef slab, PreHash pre) {
    extern __shared__ __align__(16) unsigned char s_dyn[];  // kUpdateWarps stages (may exceed 48 KB)
    UpdateStage* s_stage = reinterpret_cast<UpdateStage*>(s_dyn);
    UpdateStage& st = s_stage[threadIdx.x >> 5];
    if constexpr (!kSlab) {
        update_group<kDebug, false, kExt>(pos_rho, vel_pres, offsets, c, groups, num_groups, pos_out,
                                    vel_out, forces, list, aos_out, slab, st, blockIdx.x, gridDim.x,
                                    pre);
    } else {
        // A dead step (sticky error in the slab record) has an empty group table: the arena
        // memset zeroed the count and k_finish_sort did not build one.  (No early exit of the
        // blocks beyond the table and no test of the error word up here: with either one ptxas
        // allocates the walk loop 8 % longer; update_group's own test of the count covers both.)
        const uint32_t blocks = max(1u, (*num_groups + kUpdateWarps - 1u) / kUpdateWarps);
        update_group<kDebug, true, kExt>(pos_rho, vel_pres, offsets, c, groups, num_groups, pos_out,
                                   vel_out, forces, list, aos_out, slab, st, blockIdx.x, blocks, pre);
    }
}

struct GroupTable {
    const uint4* groups;
    const uint32_t* count;  // device scalar
    int max_groups;         // launch bound (>= *count)
};

inline int blocks_for(int groups, int warps) { return (groups + warps - 1) / warps; }

// Launchers: the instantiation is chosen from (debug outputs?, slab mode?, extended physics?).
template <bool kSlab, bool kExt>
inline void launch_density_tile_t(float4* pos_rho, float4* vel_pres, const uint32_t* offsets,
                                  const SphConstsExt& c, const GroupTable& gt,
                                  uint32_t* neighbour_counts, NbrList list, cudaStream_t stream,
                                  const SlabRef& slab) {
    const int blocks = blocks_for(gt.max_groups, kDensityWarps);
    if (neighbour_counts)
        k_density_tile<true, kSlab, kExt><<<blocks, kDensityWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, gt.groups, gt.count, neighbour_counts, list, slab);
    else
        k_density_tile<false, kSlab, kExt><<<blocks, kDensityWarps * 32, 0, stream>>>(
            pos_rho, vel_pres, offsets, c, gt.groups, gt.count, nullptr, list, slab);
}

inline void launch_density_tile(float4* pos_rho, float4* vel_pres, const uint32_t* offsets,
                                const SphConstsExt& c, const GroupTable& gt,
                                uint32_t* neighbour_counts, NbrList list, cudaStream_t stream,
                                const SlabRef& slab = SlabRef{}) {
    const bool ext = c.phys != 0u;
    if (slab.dyn && ext)
        launch_density_tile_t<true, true>(pos_rho, vel_pres, offsets, c, gt, neighbour_counts, list, stream, slab);
    else if (slab.dyn)
        launch_density_tile_t<true, false>(pos_rho, vel_pres, offsets, c, gt, neighbour_counts, list, stream, slab);
    else if (ext)
        launch_density_tile_t<false, true>(pos_rho, vel_pres, offsets, c, gt, neighbour_counts, list, stream, slab);
    else
        launch_density_tile_t<false, false>(pos_rho, vel_pres, offsets, c, gt, neighbour_counts, list, stream, slab);
}

template <bool kSlab, bool kExt>
inline void launch_update_tile_t(const float4* pos_rho, const float4* vel_pres,
                                 const uint32_t* offsets, const SphConstsExt& c, const GroupTable& gt,
                                 float4* pos_out, float4* vel_out, float4* forces, NbrList list,
                                 cudaStream_t stream, float4* aos_out, const SlabRef& slab,
                                 const PreHash& pre) {
    const int blocks = blocks_for(gt.max_groups, kUpdateWarps);
    constexpr size_t smem = kUpdateWarps * sizeof(UpdateStage);
    if (smem > 48 * 1024) {  // opt-in size; the attribute is per device, so set it per launch
        if (forces)
            cudaFuncSetAttribute(k_update_tile<true, kSlab, kExt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else
            cudaFuncSetAttribute(k_update_tile<false, kSlab, kExt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (forces)
        k_update_tile<true, kSlab, kExt><<<blocks, kUpdateWarps * 32, smem, stream>>>(
            pos_rho, vel_pres, offsets, c, gt.groups, gt.count, pos_out, vel_out, forces,
            list, aos_out, slab, pre);
    else
        k_update_tile<false, kSlab, kExt><<<blocks, kUpdateWarps * 32, smem, stream>>>(
            pos_rho, vel_pres, offsets, c, gt.groups, gt.count, pos_out, vel_out, nullptr,
            list, aos_out, slab, pre);
}

inline void launch_update_tile(const float4* pos_rho, const float4* vel_pres,
                               const uint32_t* offsets, const SphConstsExt& c, const GroupTable& gt,
                               float4* pos_out, float4* vel_out, float4* forces, NbrList list,
                               cudaStream_t stream, float4* aos_out = nullptr,
                               const SlabRef& slab = SlabRef{}, const PreHash& pre = PreHash{}) {
    const bool ext = c.phys != 0u;
    if (slab.dyn && ext)
        launch_update_tile_t<true, true>(pos_rho, vel_pres, offsets, c, gt, pos_out, vel_out, forces, list, stream, aos_out, slab, pre);
    else if (slab.dyn)
        launch_update_tile_t<true, false>(pos_rho, vel_pres, offsets, c, gt, pos_out, vel_out, forces, list, stream, aos_out, slab, pre);
    else if (ext)
        launch_update_tile_t<false, true>(pos_rho, vel_pres, offsets, c, gt, pos_out, vel_out, forces, list, stream, aos_out, slab, pre);
    else
        launch_update_tile_t<false, false>(pos_rho, vel_pres, offsets, c, gt, pos_out, vel_out, forces, list, stream, aos_out, slab, pre);
}

}  // namespace wc
