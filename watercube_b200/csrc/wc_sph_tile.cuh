// wc_sph_tile.cuh -- warp-cooperative shared-memory gather kernels (placeholder until the
// tiled path lands; the launchers return -1 = "not covered", so the simple path runs).
#pragma once

#include "wc_common.cuh"

namespace wc {

inline int launch_density_tile(float4*, float4*, const uint32_t*, const SphConsts&, uint32_t*,
                               cudaStream_t) {
    return -1;
}

inline int launch_update_tile(const float4*, const float4*, const uint32_t*, const SphConsts&,
                              float4*, float4*, float4*, cudaStream_t) {
    return -1;
}

}  // namespace wc
