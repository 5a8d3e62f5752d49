// wc_common.cuh -- shared device helpers and the per-launch constant block.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wc {

constexpr int kWarp = 32;

// Everything the density / update kernels read as GLSL uniforms
// (src/core/Fluid.cpp:276-285, :305-317), plus values precomputed once on the host.
struct SphConsts {
    int n;           // numParticles (targets of this launch)
    int first;       // index of the first target in the candidate arrays (0 unless slab mode)
    int G;           // gridRes
    int Gz;          // z-layers covered by the offsets table (G, or slab layers + 2 ghost layers)
    int zbase;       // global z-layer of the table's layer 0 (0, or slab z_begin - 1)
    float bin;       // binSize
    float size;      // size
    float h;         // kernelRadius
    float h2;        // h*h (fp32 product, as poly6Kernel forms it)
    float T;         // smallest fp32 x with sqrtf(x) >= h: "dist >= h" <=> d2 >= T
    float m;         // particleMass
    float poly6C, spikyC, viscC;
    float mu, k, rho0, P0;
    float g[3];
    float dt;
    float mo[3], md[3];
    int mouse_hits;  // update.comp:118-121 evaluated on the host (all-uniform expression)
};

// Slab mode with attached neighbours (wc_slab_peer_*): where this rank's first / last owned
// layer lives in the neighbours' buffer 2 (their ghost slots, mapped peer memory).  The kernels
// that PRODUCE halo data store it there as they go -- the reorder writes positions and
// velocities, the density pass adds density and pressure -- so the halo crosses NVLink under
// the producing kernel and no copy sits between the phases.  All null / zero otherwise.
struct PeerHalo {
    float4* pos[2];     // [0] = the rank below, [1] = the rank above
    float4* vel[2];
    uint32_t dst[2];    // index of this rank's first halo particle in that neighbour's arrays
    uint32_t n_first;   // owned sorted particles [0, n_first) are the lower halo
    uint32_t hi_begin;  // owned sorted particles [hi_begin, n) are the upper halo
};

// count.comp:32, one component: clamp(int(p / binSize), 0, gridRes - 1) with an IEEE
// fp32 divide and truncation toward zero.  Same float-side clamp as the oracle
// (oracle/wc_oracle.cpp cell_coord) so NaN / huge inputs are defined identically.
__device__ __forceinline__ int cell_coord(float p, float bin, int G) {
    const float q = __fdiv_rn(p, bin);
    if (!(q >= 1.0f)) return 0;
    if (q >= (float)G) return G - 1;
    return __float2int_rz(q);
}

// count.comp:33.  zbase shifts the z-layer for a slab-local table (0 for the whole grid).
__device__ __forceinline__ uint32_t cell_index(float x, float y, float z, float bin, int G,
                                               int zbase = 0) {
    const uint32_t cx = (uint32_t)cell_coord(x, bin, G);
    const uint32_t cy = (uint32_t)cell_coord(y, bin, G);
    const uint32_t cz = (uint32_t)(cell_coord(z, bin, G) - zbase);
    return (cz * (uint32_t)G + cy) * (uint32_t)G + cx;
}

// Squared distance with the op order the oracle pins: fma(rz,rz, fma(ry,ry, rx*rx)).
__device__ __forceinline__ float dist2(float rx, float ry, float rz) {
    return __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, __fmul_rn(rx, rx)));
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace wc
