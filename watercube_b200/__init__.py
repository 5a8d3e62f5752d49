"""watercube_b200 -- B200-native SPH step (the WaterCube hot path) behind a C-ABI.

The native library (watercube_b200/csrc/libwc_sph.so) is loaded lazily by
``watercube_b200.capi``; there is no CPU fallback: every compute entry point raises
if the library or a CUDA device is missing.
"""
__all__ = ["capi", "scenes"]
