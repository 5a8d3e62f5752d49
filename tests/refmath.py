"""Independent restatements used to pin the C++ oracle (tests only).

* ``brute_*``: numpy all-pairs O(N^2) fp32 evaluation of density.comp / update.comp
  with no grid at all (valid because binSize >= kernelRadius, so every particle
  within h lies in the 27-cell neighbourhood).
* ``closed_form_*``: float64 formulas for one- and two-particle configurations.
"""
import numpy as np

f32 = np.float32


def fma32(a, b, c):
    """fp32 fma emulated in fp64 (exact product, one fp64 add, round to fp32)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def pair_dist(pos):
    r = pos[:, None, :] - pos[None, :, :]  # fp32, r[i,j] = p_i - p_j
    d2 = fma32(r[..., 2], r[..., 2], fma32(r[..., 1], r[..., 1], r[..., 0] * r[..., 0]))
    return r, np.sqrt(d2)  # np.sqrt on float32 is correctly rounded


def consts(params, derived):
    return dict(h=f32(derived.kernel_radius), m=f32(derived.particle_mass),
                poly6C=f32(derived.poly6_const), spikyC=f32(derived.spiky_const),
                viscC=f32(derived.visc_const), size=f32(params.size),
                k=f32(params.stiffness), rho0=f32(params.rest_density),
                P0=f32(params.rest_pressure), mu=f32(params.viscosity_coefficient),
                g=np.array(list(params.gravity), f32))


def poly6(c, r):
    t = c["h"] * c["h"] - r * r
    return ((t * t) * t) * c["poly6C"]


def wall_density(c, pos):
    h, m, size = c["h"], c["m"], c["size"]
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    d = np.zeros(len(pos), f32)
    d += np.where(x < h, m * poly6(c, x), np.where(x > size - h, m * poly6(c, size - x), f32(0)))
    d += np.where(y < h, m * poly6(c, y), np.where(y > size - h, m * poly6(c, size - y), f32(0)))
    # density.comp:72-76: the z-high branch tests p.y (quirk Q2)
    d += np.where(z < h, m * poly6(c, z), np.where(y > size - h, m * poly6(c, size - z), f32(0)))
    return d * f32(4)


def brute_density(pos, params, derived):
    """-> (density_with_wall, pressure, neighbour_counts); fp32, summed in index order."""
    c = consts(params, derived)
    n = len(pos)
    _, dist = pair_dist(pos)
    mask = (dist < c["h"]) & ~np.eye(n, dtype=bool)
    w = np.where(mask, c["m"] * poly6(c, dist), f32(0)).astype(f32)
    rho = np.full(n, c["m"] * poly6(c, f32(0)), f32)
    for j in range(n):  # sequential fp32 accumulation (order differs from the oracle's)
        rho = rho + w[:, j]
    q = rho / c["rho0"]
    pres = c["P0"] + c["k"] * (((q * q) * q) - f32(1))
    return rho + wall_density(c, pos), pres, mask.sum(1).astype(np.uint32)


def brute_forces_f64(pos, vel, rho, pres, params, derived):
    """Pressure + viscosity + gravity force in float64 from fp32 inputs (no wall/mouse)."""
    c = {k: (np.float64(v) if np.ndim(v) == 0 else v.astype(np.float64))
         for k, v in consts(params, derived).items()}
    n = len(pos)
    _, dist32 = pair_dist(pos)
    mask = (dist32 < f32(derived.kernel_radius)) & ~np.eye(n, dtype=bool)
    p64, v64 = pos.astype(np.float64), vel.astype(np.float64)
    rho, pres = rho.astype(np.float64), pres.astype(np.float64)
    r = p64[:, None, :] - p64[None, :, :]
    d = np.sqrt((r * r).sum(-1))
    d_safe = np.where(d > 0, d, 1.0)
    pr = (pres[:, None] + pres[None, :]) / (2.0 * rho[None, :])
    s = (c["h"] - d) ** 2 * c["spikyC"]
    wp = np.where((mask & (pr > 0) & (d > 0))[..., None],
                  (c["m"] * pr * s / d_safe)[..., None] * r, 0.0)
    Fp = -wp.sum(1)
    wv = np.where(mask, (c["h"] - d) * c["viscC"], 0.0)
    Fv = (c["m"] * (v64[None, :, :] - v64[:, None, :]) / rho[None, :, None] * wv[..., None]).sum(1)
    return Fp + c["mu"] * Fv + c["g"][None, :] * rho[:, None]
