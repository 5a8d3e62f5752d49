"""Synthetic scenes for the SPH hot path.

The reference seeds libc ``rand()`` (src/core/Fluid.cpp:105,114), which is not
portable (quirk Q16), so scenes use the counter-based generator below.  The C++
facade (csrc/core/Fluid.cpp) implements the very same generator, so a scene built
there equals the one built here bit for bit.

Generator (stated in full because results depend on it):
    x  = (counter + seed * 0x9E3779B9) mod 2^32
    x ^= x >> 16;  x *= 0x7FEB352D;  x ^= x >> 15;  x *= 0x846CA68B;  x ^= x >> 16
    u  = float32(x >> 8) * 2^-24                      in [0, 1)
with counter = 3 * particle_index + component for positions/jitter and
counter = 3 * (N + particle_index) + component for velocities.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PARTICLE_FLOATS = 8  # x y z density | vx vy vz pressure  (src/core/util.h:29-35)

# Fluid::Fluid defaults, src/core/Fluid.cpp:9-27
DEFAULT_RADIUS = 0.01
DEFAULT_SPACING_FACTOR = 1.75  # Fluid.cpp:110
FILL_RATIO = 0.77              # 44 * 0.0175 / 1.0 for the default scene
CELLS_PER_UNIT = 21            # default gridRes / size


def hash_u01(counter: np.ndarray, seed: int) -> np.ndarray:
    """Counter-based uniform float32 in [0,1) (see module docstring)."""
    x = (np.asarray(counter, dtype=np.uint64) + np.uint64((seed * 0x9E3779B9) & 0xFFFFFFFF)) \
        & np.uint64(0xFFFFFFFF)
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def lattice_side(n: int) -> int:
    """int(ceil(cbrt(n))) of Fluid.cpp:111, computed in integers."""
    d = int(round(n ** (1.0 / 3.0)))
    while d * d * d < n:
        d += 1
    while d > 0 and (d - 1) ** 3 >= n:
        d -= 1
    return d


@dataclass
class Scene:
    name: str
    particles: np.ndarray  # [n, 8] float32 AoS
    size: float
    grid_res: int
    particle_radius: float

    @property
    def n(self) -> int:
        return int(self.particles.shape[0])


def scaled_box(n: int, radius: float = DEFAULT_RADIUS):
    """Box size / gridRes that keep the default scene's fill ratio and bin/h ratio
    (scale the box, not the radius).  -> (size, grid_res)"""
    if n == 80000 and radius == DEFAULT_RADIUS:
        return 1.0, 21  # Fluid.cpp:10,14
    d = lattice_side(n)
    block = d * (radius * DEFAULT_SPACING_FACTOR)
    size = float(np.float32(block / FILL_RATIO))
    grid_res = int(np.floor(size * CELLS_PER_UNIT * (DEFAULT_RADIUS / radius)))
    return size, max(grid_res, 1)


def dam_break(n: int, seed: int = 0, radius: float = DEFAULT_RADIUS, size: float | None = None,
              grid_res: int | None = None, chunk: int = 1 << 22) -> Scene:
    """Fluid::generateInitialParticles (src/core/Fluid.cpp:104-134): jittered lattice
    block in the origin corner, index = z*d*d + y*d + x, zero velocity."""
    if size is None or grid_res is None:
        s, g = scaled_box(n, radius)
        size = s if size is None else size
        grid_res = g if grid_res is None else grid_res
    d = lattice_side(n)
    distance = np.float32(radius) * np.float32(DEFAULT_SPACING_FACTOR)  # Fluid.cpp:110
    jitter = distance * np.float32(0.5)                                 # Fluid.cpp:113
    half = jitter / np.float32(2.0)
    out = np.zeros((n, PARTICLE_FLOATS), np.float32)
    for s0 in range(0, n, chunk):
        s1 = min(n, s0 + chunk)
        idx = np.arange(s0, s1, dtype=np.int64)
        xyz = np.stack([idx % d, (idx // d) % d, idx // (d * d)], axis=1).astype(np.float32)
        pos = xyz * distance                                            # Fluid.cpp:129
        for k in range(3):
            u = hash_u01(3 * idx + k, seed)
            pos[:, k] += u * jitter - half                              # Fluid.cpp:114,130
        out[s0:s1, 0:3] = pos
    return Scene(f"dam_break_{n}", out, float(size), int(grid_res), float(radius))


def uniform_box(n: int, size: float = 3.5, h: float = 0.04, seed: int = 0,
                chunk: int = 1 << 22) -> Scene:
    """Uniform random box (the commented-out init at Fluid.cpp:124-126; BASELINE.md
    config 5): x ~ U[0.001, size-0.001]^3, v ~ U[-1,1]^3, particleRadius = h/4,
    gridRes = floor(size/h)."""
    out = np.zeros((n, PARTICLE_FLOATS), np.float32)
    lo = np.float32(0.001)
    span = np.float32(size) - np.float32(0.002)
    for s0 in range(0, n, chunk):
        s1 = min(n, s0 + chunk)
        idx = np.arange(s0, s1, dtype=np.int64)
        for k in range(3):
            out[s0:s1, k] = lo + hash_u01(3 * idx + k, seed) * span
            out[s0:s1, 4 + k] = hash_u01(3 * (n + idx) + k, seed) * np.float32(2.0) - np.float32(1.0)
    grid_res = max(int(np.floor(size / h)), 1)
    return Scene(f"uniform_box_{n}_h{h:g}", out, float(size), grid_res, float(h) / 4.0)


def smoothing_length_for_neighbours(nb: float, number_density: float = 1.0 / 0.0175 ** 3) -> float:
    """h = (3 nb / (4 pi n))^(1/3)  (BASELINE.md config 5)."""
    return float((3.0 * nb / (4.0 * np.pi * number_density)) ** (1.0 / 3.0))
