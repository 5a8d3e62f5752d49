// wc_diag.cuh -- on-device inspection of a particle buffer and of the cell table.
//
// The reference's only verification tooling is visual: render modes 1-4 of particle.vert:35-55
// colour a particle red when it left the box or its density is not positive, and grid.vert:29-43
// / Sort::printGrids (Sort.cpp:237-249) / util::printParticles (util.cpp:113-126) dump the
// count / offset tables and particles to the debug console.  A headless solver needs the same
// facts as numbers: how many particles are invalid, the conserved quantities, the density
// distribution and the cell occupancy -- computed where the data lives (one pass over 32 B per
// particle) instead of downloading the buffers.
//
// Deterministic: every block reduces a fixed slice in fp64 and a single block folds the
// per-block partials in index order, so the same buffer always gives the same bits.
#pragma once

#include "../../include/wc_sph.h"
#include "wc_common.cuh"

namespace wc {

constexpr int kDiagThreads = 256;
constexpr int kDiagMaxBlocks = 1184;  // 148 SMs x 8 resident blocks
constexpr float kMaxSpeed = 50.0f;    // update.comp:5

// What one block (and, after the fold, the whole launch) knows.  Sums are over VALID
// particles only, so one NaN does not erase the statistics it is reported next to.
struct DiagPartial {
    double mom[3], ke, com[3], rho_sum, pres_sum;
    float vmax2, rho_min, rho_max, pres_min, pres_max;
    unsigned long long valid, invalid, out_of_box, at_clamp;
    unsigned long long hist[WC_DIAG_HIST_BINS];
};

__device__ __forceinline__ void diag_identity(DiagPartial& a) {
    for (int k = 0; k < 3; k++) a.mom[k] = 0.0, a.com[k] = 0.0;
    a.ke = a.rho_sum = a.pres_sum = 0.0;
    a.vmax2 = 0.0f;
    a.rho_min = a.pres_min = INFINITY;
    a.rho_max = a.pres_max = -INFINITY;
    a.valid = a.invalid = a.out_of_box = a.at_clamp = 0ull;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Scalars of a partial as one flat list, so both reduction levels share the code.
constexpr int kDiagF64 = 9, kDiagU64 = 4;
__device__ __forceinline__ double& diag_f64(DiagPartial& a, int k) {
    return k < 3 ? a.mom[k] : k == 3 ? a.ke : k < 7 ? a.com[k - 4] : k == 7 ? a.rho_sum : a.pres_sum;
}
__device__ __forceinline__ unsigned long long& diag_u64(DiagPartial& a, int k) {
    return k == 0 ? a.valid : k == 1 ? a.invalid : k == 2 ? a.out_of_box : a.at_clamp;
}

// Block-wide fold of per-thread partials (histogram excluded); the result is valid in thread 0.
__device__ __forceinline__ void diag_block_fold(DiagPartial& a, DiagPartial* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kDiagF64; k++) diag_f64(a, k) = warp_sum_f64(diag_f64(a, k));
#pragma unroll
    for (int k = 0; k < kDiagU64; k++) diag_u64(a, k) = warp_sum_u64(diag_u64(a, k));
    a.vmax2 = warp_max_f32(a.vmax2);
    a.rho_min = warp_min_f32(a.rho_min), a.rho_max = warp_max_f32(a.rho_max);
    a.pres_min = warp_min_f32(a.pres_min), a.pres_max = warp_max_f32(a.pres_max);
    if (lane == 0) {
        DiagPartial& w = s_warp[warp];
        for (int k = 0; k < kDiagF64; k++) diag_f64(w, k) = diag_f64(a, k);
        for (int k = 0; k < kDiagU64; k++) diag_u64(w, k) = diag_u64(a, k);
        w.vmax2 = a.vmax2;
        w.rho_min = a.rho_min, w.rho_max = a.rho_max;
        w.pres_min = a.pres_min, w.pres_max = a.pres_max;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < kDiagThreads / 32; q++) {  // warp order: deterministic
            DiagPartial& w = s_warp[q];
            for (int k = 0; k < kDiagF64; k++) diag_f64(a, k) += diag_f64(w, k);
            for (int k = 0; k < kDiagU64; k++) diag_u64(a, k) += diag_u64(w, k);
            a.vmax2 = fmaxf(a.vmax2, w.vmax2);
            a.rho_min = fminf(a.rho_min, w.rho_min), a.rho_max = fmaxf(a.rho_max, w.rho_max);
            a.pres_min = fminf(a.pres_min, w.pres_min), a.pres_max = fmaxf(a.pres_max, w.pres_max);
        }
    }
}

// Pass 1: block b covers particles [b * per_block, (b + 1) * per_block).
__global__ void __launch_bounds__(kDiagThreads)
k_diag_particles(const float4* __restrict__ pos_rho, const float4* __restrict__ vel_pres, int n,
                 int per_block, float size, float inv_rho0, DiagPartial* __restrict__ partials) {
    __shared__ DiagPartial s_warp[kDiagThreads / 32];
    __shared__ unsigned int s_hist[WC_DIAG_HIST_BINS];
    if (threadIdx.x < WC_DIAG_HIST_BINS) s_hist[threadIdx.x] = 0u;
    __syncthreads();
    DiagPartial a;
    diag_identity(a);
    const int begin = blockIdx.x * per_block, end = min(n, begin + per_block);
    for (int i = begin + threadIdx.x; i < end; i += kDiagThreads) {
        const float4 p = __ldg(&pos_rho[i]), v = __ldg(&vel_pres[i]);
        // particle.vert:36-46 (render modes 1 / 2): outside the box, or density not positive
        const bool inside = p.x >= 0.0f && p.x <= size && p.y >= 0.0f && p.y <= size &&
                            p.z >= 0.0f && p.z <= size;  // false for NaN
        const float v2 = v.x * v.x + v.y * v.y + v.z * v.z;
        const bool finite = isfinite(v2) && isfinite(p.w) && isfinite(v.w);
        if (!inside) a.out_of_box++;
        if (!inside || !finite || !(p.w > 0.0f)) {
            a.invalid++;
            continue;
        }
        a.valid++;
        a.mom[0] += (double)v.x, a.mom[1] += (double)v.y, a.mom[2] += (double)v.z;
        a.ke += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z;
        a.com[0] += (double)p.x, a.com[1] += (double)p.y, a.com[2] += (double)p.z;
        a.rho_sum += (double)p.w, a.pres_sum += (double)v.w;
        a.vmax2 = fmaxf(a.vmax2, v2);
        a.rho_min = fminf(a.rho_min, p.w), a.rho_max = fmaxf(a.rho_max, p.w);
        a.pres_min = fminf(a.pres_min, v.w), a.pres_max = fmaxf(a.pres_max, v.w);
        if (fabsf(v.x) >= kMaxSpeed || fabsf(v.y) >= kMaxSpeed || fabsf(v.z) >= kMaxSpeed)
            a.at_clamp++;  // update.comp:199 clamped at least one component
        const float r = p.w * inv_rho0 * (float)WC_DIAG_HIST_PER_UNIT;
        const int bin = r >= (float)(WC_DIAG_HIST_BINS - 1) ? WC_DIAG_HIST_BINS - 1 : (int)r;
        atomicAdd(&s_hist[bin], 1u);
    }
    diag_block_fold(a, s_warp);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < WC_DIAG_HIST_BINS; k++) a.hist[k] = s_hist[k];
        partials[blockIdx.x] = a;
    }
}

// Pass 2: one block folds the partials in index order.
__global__ void __launch_bounds__(kDiagThreads)
k_diag_fold(const DiagPartial* __restrict__ partials, int blocks, DiagPartial* __restrict__ out) {
    __shared__ DiagPartial s_warp[kDiagThreads / 32];
    __shared__ unsigned long long s_hist[WC_DIAG_HIST_BINS];
    if (threadIdx.x < WC_DIAG_HIST_BINS) {
        unsigned long long t = 0ull;
        for (int b = 0; b < blocks; b++) t += partials[b].hist[threadIdx.x];
        s_hist[threadIdx.x] = t;
    }
    DiagPartial a;
    diag_identity(a);
    // thread t owns blocks t, t + 256, ... : a fixed assignment, folded in a fixed order
    for (int b = threadIdx.x; b < blocks; b += kDiagThreads) {
        const DiagPartial& w = partials[b];
        for (int k = 0; k < 3; k++) a.mom[k] += w.mom[k], a.com[k] += w.com[k];
        a.ke += w.ke, a.rho_sum += w.rho_sum, a.pres_sum += w.pres_sum;
        a.valid += w.valid, a.invalid += w.invalid, a.out_of_box += w.out_of_box;
        a.at_clamp += w.at_clamp;
        a.vmax2 = fmaxf(a.vmax2, w.vmax2);
        a.rho_min = fminf(a.rho_min, w.rho_min), a.rho_max = fmaxf(a.rho_max, w.rho_max);
        a.pres_min = fminf(a.pres_min, w.pres_min), a.pres_max = fmaxf(a.pres_max, w.pres_max);
    }
    diag_block_fold(a, s_warp);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < WC_DIAG_HIST_BINS; k++) a.hist[k] = s_hist[k];
        *out = a;
    }
}

// Cell occupancy from the offsets table of the last sort (what Sort::printGrids prints):
// out[0] = largest cell count, out[1] = non-empty cells.  Integer atomics: order-free.
__global__ void k_diag_cells(const uint32_t* __restrict__ offsets, int num_bins,
                             unsigned int* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int cnt = 0u;
    if (c < num_bins) cnt = offsets[c + 1] - offsets[c];
    const unsigned int wmax = __reduce_max_sync(0xffffffffu, cnt);
    const unsigned int wnz = __popc(__ballot_sync(0xffffffffu, cnt != 0u));
    if ((threadIdx.x & 31) == 0 && wnz) {
        atomicMax(&out[0], wmax);
        atomicAdd(&out[1], wnz);
    }
}

}  // namespace wc
