"""ctypes binding of the CPU oracle (oracle/libwc_oracle.so).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by watercube_b200/ (the product path has no CPU
fallback).  See oracle/wc_oracle.h for what each entry point restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwc_oracle.so")

PARTICLE_DTYPE = np.dtype(
    [("position", np.float32, 3), ("density", np.float32), ("velocity", np.float32, 3),
     ("pressure", np.float32)]
)
assert PARTICLE_DTYPE.itemsize == 32


class Params(C.Structure):
    _fields_ = [
        ("num_particles", C.c_int32),
        ("grid_res", C.c_int32),
        ("size", C.c_float),
        ("particle_radius", C.c_float),
        ("viscosity_coefficient", C.c_float),
        ("stiffness", C.c_float),
        ("rest_density", C.c_float),
        ("rest_pressure", C.c_float),
        ("gravity", C.c_float * 3),
        ("time_scale", C.c_float),
        ("mouse_origin", C.c_float * 3),
        ("mouse_dir", C.c_float * 3),
        # extended physics (wc_oracle.h WCO_PHYS_*); 0 = the reference's step
        ("physics_flags", C.c_uint32),
        ("surface_tension", C.c_float),
        ("surface_threshold", C.c_float),
        ("wall_stiffness", C.c_float),
        ("wall_distance", C.c_float),
        ("wall_rest_density", C.c_float),
    ]


PHYS_WALL_PARTICLES = 1
PHYS_SURFACE_TENSION = 2


class Derived(C.Structure):
    _fields_ = [
        ("num_bins", C.c_int32),
        ("bin_size", C.c_float),
        ("kernel_radius", C.c_float),
        ("particle_mass", C.c_float),
        ("poly6_const", C.c_float),
        ("spiky_const", C.c_float),
        ("visc_const", C.c_float),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only; no reference sources involved)."""
    src = os.path.join(_HERE, "wc_oracle.cpp")
    hdr = os.path.join(_HERE, "wc_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libwc_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.wco_default_params.argtypes = [C.POINTER(Params)]
        L.wco_derive.argtypes = [C.POINTER(Params), C.POINTER(Derived)]
        L.wco_cell_ids.argtypes = [vp, i32, f32, i32, vp]
        L.wco_sort.argtypes = [vp, i32, f32, i32, vp, vp, vp, vp, vp]
        L.wco_density.argtypes = [vp, i32, vp, vp, C.POINTER(Params), vp, i32]
        L.wco_update.argtypes = [vp, vp, i32, vp, vp, C.POINTER(Params), f32, vp, i32]
        L.wco_step.argtypes = [vp, vp, i32, C.POINTER(Params), f32, vp, vp, i32]
        L.wco_advect.argtypes = [vp, i32, f32, f32]
        L.wco_density_f64.argtypes = [vp, i32, vp, vp, C.POINTER(Params), vp, vp, i32]
        L.wco_update_f64.argtypes = [vp, vp, vp, i32, vp, vp, C.POINTER(Params), f32, vp, vp, vp,
                                     i32]
        L.wco_update_f64_scaled.argtypes = [vp, vp, vp, i32, vp, vp, C.POINTER(Params), f32, vp, vp,
                                            vp, vp, i32]
        L.wco_update_f64_scaled.restype = None
        L.wco_max_threads.restype = i32
        for f in ("wco_default_params", "wco_derive", "wco_cell_ids", "wco_sort", "wco_density",
                  "wco_update", "wco_step", "wco_advect", "wco_density_f64", "wco_update_f64"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(**overrides) -> Params:
    p = Params()
    lib().wco_default_params(C.byref(p))
    for k, v in overrides.items():
        if k in ("gravity", "mouse_origin", "mouse_dir"):
            getattr(p, k)[:] = [float(x) for x in v]
        else:
            setattr(p, k, v)
    return p


def derive(p: Params) -> Derived:
    d = Derived()
    lib().wco_derive(C.byref(p), C.byref(d))
    return d


def max_threads() -> int:
    return int(lib().wco_max_threads())


def as_particles(a) -> np.ndarray:
    """Accept [n,8] float32 or structured; return a C-contiguous structured view/copy."""
    a = np.ascontiguousarray(a)
    if a.dtype == PARTICLE_DTYPE:
        return a
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 8)
    return a.view(PARTICLE_DTYPE).reshape(-1)


def as_f32(a: np.ndarray) -> np.ndarray:
    """Structured particles -> [n,8] float32 view."""
    return a.view(np.float32).reshape(-1, 8)


def cell_ids(particles, bin_size, grid_res) -> np.ndarray:
    P = as_particles(particles)
    out = np.empty(P.shape[0], np.uint32)
    lib().wco_cell_ids(_ptr(P), P.shape[0], float(bin_size), int(grid_res), _ptr(out))
    return out


def sort(particles, bin_size, grid_res):
    """-> dict(cell_ids, counts, offsets, perm, sorted)"""
    P = as_particles(particles)
    n = P.shape[0]
    B = int(grid_res) ** 3
    out = dict(
        cell_ids=np.empty(n, np.uint32),
        counts=np.empty(B, np.uint32),
        offsets=np.empty(B, np.uint32),
        perm=np.empty(n, np.uint32),
        sorted=np.empty(n, PARTICLE_DTYPE),
    )
    lib().wco_sort(_ptr(P), n, float(bin_size), int(grid_res), _ptr(out["cell_ids"]),
                   _ptr(out["counts"]), _ptr(out["offsets"]), _ptr(out["perm"]),
                   _ptr(out["sorted"]))
    return out


def density(sorted_particles, counts, offsets, params: Params, nthreads=1):
    """In-place density/pressure on a COPY; -> (particles, neighbour_counts)"""
    P = as_particles(sorted_particles).copy()
    nc = np.empty(P.shape[0], np.uint32)
    lib().wco_density(_ptr(P), P.shape[0], _ptr(counts), _ptr(offsets), C.byref(params), _ptr(nc),
                      int(nthreads))
    return P, nc


def update(in_particles, counts, offsets, params: Params, dt, nthreads=1):
    """-> (out_particles, forces[n,3])"""
    P = as_particles(in_particles)
    out = np.empty_like(P)
    F = np.empty((P.shape[0], 3), np.float32)
    lib().wco_update(_ptr(P), _ptr(out), P.shape[0], _ptr(counts), _ptr(offsets), C.byref(params),
                     float(dt), _ptr(F), int(nthreads))
    return out, F


class Stepper:
    """Fluid::update loop on the CPU (Fluid.cpp:342-354) with persistent scratch."""

    def __init__(self, particles, params: Params, nthreads=1):
        self.params = params
        self.n = int(params.num_particles)
        self.buf1 = as_particles(particles).copy()
        assert self.buf1.shape[0] == self.n
        self.buf2 = np.empty_like(self.buf1)
        B = int(params.grid_res) ** 3
        self.counts = np.empty(B, np.uint32)
        self.offsets = np.empty(B, np.uint32)
        self.nthreads = int(nthreads)

    def step(self, frame_dt=1.0 / 60.0):
        lib().wco_step(_ptr(self.buf1), _ptr(self.buf2), self.n, C.byref(self.params),
                       float(frame_dt), _ptr(self.counts), _ptr(self.offsets), self.nthreads)


def density_f64(sorted_particles, counts, offsets, params: Params, nthreads=1):
    P = as_particles(sorted_particles)
    rho = np.empty(P.shape[0], np.float64)
    pres = np.empty(P.shape[0], np.float64)
    lib().wco_density_f64(_ptr(P), P.shape[0], _ptr(counts), _ptr(offsets), C.byref(params),
                          _ptr(rho), _ptr(pres), int(nthreads))
    return rho, pres


def update_f64(in_particles, rho, pres, counts, offsets, params: Params, dt, nthreads=1):
    P = as_particles(in_particles)
    n = P.shape[0]
    F = np.empty((n, 3), np.float64)
    v = np.empty((n, 3), np.float64)
    x = np.empty((n, 3), np.float64)
    lib().wco_update_f64(_ptr(P), _ptr(rho), _ptr(pres), n, _ptr(counts), _ptr(offsets),
                         C.byref(params), float(dt), _ptr(F), _ptr(v), _ptr(x), int(nthreads))
    return F, v, x


def update_f64_scaled(in_particles, rho, pres, counts, offsets, params: Params, dt, nthreads=1):
    """update_f64 plus, per particle and component, the sum of the magnitudes of all terms
    added into F (the scale rounding errors are proportional to).  -> (F, v, x, scale)"""
    P = as_particles(in_particles)
    n = P.shape[0]
    F = np.empty((n, 3), np.float64)
    v = np.empty((n, 3), np.float64)
    x = np.empty((n, 3), np.float64)
    S = np.empty((n, 3), np.float64)
    lib().wco_update_f64_scaled(_ptr(P), _ptr(rho), _ptr(pres), n, _ptr(counts), _ptr(offsets),
                                C.byref(params), float(dt), _ptr(F), _ptr(v), _ptr(x), _ptr(S),
                                int(nthreads))
    return F, v, x, S


def host_threads() -> int:
    """Processors this process may run on (not OMP_NUM_THREADS, which torchrun sets to 1)."""
    import os

    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def advect(particles, size, dt):
    P = as_particles(particles).copy()
    lib().wco_advect(_ptr(P), P.shape[0], float(size), float(dt))
    return P
