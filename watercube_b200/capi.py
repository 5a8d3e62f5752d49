"""ctypes binding of include/wc_sph.h plus a thin Python mirror of core::Fluid.

This is glue for tests and bench.py; the product is the native library.  There is no CPU
fallback: a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# WC_SPH_LIB selects another build of the same library (kernel-tuning experiments only).
LIB_PATH = os.environ.get("WC_SPH_LIB") or os.path.join(_CSRC, "libwc_sph.so")

WC_OK, WC_ERR_INVALID, WC_ERR_NO_DEVICE, WC_ERR_CUDA, WC_ERR_CAPACITY = 0, 1, 2, 3, 4
FLAG_DEBUG_OUTPUTS, FLAG_STAGE_TIMING, FLAG_SIMPLE_KERNELS = 1, 2, 4
STAGES = ("hash_count", "scan", "reorder", "density", "update")
NUM_STAGES = len(STAGES)

PARTICLE_DTYPE = np.dtype(
    [("position", np.float32, 3), ("density", np.float32), ("velocity", np.float32, 3),
     ("pressure", np.float32)]
)

# Every symbol include/wc_sph.h declares (checked by tests/test_capi_cpu.py).
EXPORTS = (
    "wc_abi_version", "wc_last_error", "wc_default_params", "wc_default_step_params", "wc_derive",
    "wc_device_count", "wc_create", "wc_destroy", "wc_get_derived", "wc_upload_particles", "wc_download_particles",
    "wc_step", "wc_step_host", "wc_step_export", "wc_sort_only", "wc_density_only", "wc_update_only", "wc_download_cells",
    "wc_download_forces", "wc_upload_sorted", "wc_device_ptrs", "wc_export_aos_device", "wc_sync",
    "wc_stage_times", "wc_launch_count", "wc_slab_get_view", "wc_slab_clear_recv",
    "wc_slab_sort_count", "wc_slab_sync_info", "wc_slab_reorder", "wc_slab_density",
    "wc_slab_update", "wc_advect_only", "wc_slab_ipc_export", "wc_slab_peer_open",
    "wc_slab_peer_attach", "wc_diagnose", "wc_slab_step_peer", "wc_slab_step_peer_host",
    "wc_default_physics", "wc_set_physics", "wc_get_physics", "wc_get_num_particles",
)


class Params(C.Structure):
    _fields_ = [
        ("num_particles", C.c_int32), ("capacity", C.c_int32), ("grid_res", C.c_int32),
        ("size", C.c_float), ("particle_radius", C.c_float), ("time_scale", C.c_float),
        ("device", C.c_int32), ("flags", C.c_uint32), ("neighbour_list_words", C.c_int32),
        ("slab_z_begin", C.c_int32), ("slab_z_end", C.c_int32),
        ("slab_ghost_capacity", C.c_int32), ("slab_migrant_capacity", C.c_int32),
        ("stream", C.c_void_p),
    ]


class StepParams(C.Structure):
    _fields_ = [
        ("viscosity_coefficient", C.c_float), ("stiffness", C.c_float),
        ("rest_density", C.c_float), ("rest_pressure", C.c_float),
        ("gravity", C.c_float * 3), ("mouse_origin", C.c_float * 3), ("mouse_dir", C.c_float * 3),
    ]


PHYS_WALL_PARTICLES = 1
PHYS_SURFACE_TENSION = 2


class Physics(C.Structure):
    """wc_physics: the report's future-work physics (flags 0 = the reference's step)."""
    _fields_ = [
        ("flags", C.c_uint32), ("surface_tension", C.c_float), ("surface_threshold", C.c_float),
        ("wall_stiffness", C.c_float), ("wall_distance", C.c_float),
        ("wall_rest_density", C.c_float),
    ]


class Derived(C.Structure):
    _fields_ = [
        ("num_bins", C.c_int32), ("bin_size", C.c_float), ("kernel_radius", C.c_float),
        ("particle_mass", C.c_float), ("poly6_const", C.c_float), ("spiky_const", C.c_float),
        ("visc_const", C.c_float), ("dist2_threshold", C.c_float),
    ]


class DeviceView(C.Structure):
    _fields_ = [
        ("pos_rho", C.c_void_p * 2), ("vel_pres", C.c_void_p * 2), ("cell_ids", C.c_void_p),
        ("counts", C.c_void_p), ("offsets", C.c_void_p), ("sorted", C.c_void_p),
        ("neighbour_counts", C.c_void_p), ("forces", C.c_void_p), ("stream", C.c_void_p),
        ("num_particles", C.c_int32), ("capacity", C.c_int32),
    ]


class SlabView(C.Structure):
    _fields_ = [
        ("mig_out", C.c_void_p * 2), ("mig_in", C.c_void_p * 2), ("lc_send", C.c_void_p * 2),
        ("lc_recv", C.c_void_p * 2), ("pos_rho_sorted", C.c_void_p),
        ("vel_pres_sorted", C.c_void_p), ("mig_bytes", C.c_uint64), ("lc_bytes", C.c_uint64),
        ("owned_first", C.c_int32), ("reserved", C.c_int32),
    ]


DIAG_HIST_BINS, DIAG_HIST_PER_UNIT = 32, 4


class Diagnostics(C.Structure):
    _fields_ = [
        ("particles", C.c_int64), ("invalid", C.c_int64), ("out_of_box", C.c_int64),
        ("at_speed_clamp", C.c_int64), ("mass", C.c_double), ("momentum", C.c_double * 3),
        ("kinetic_energy", C.c_double), ("centre_of_mass", C.c_double * 3),
        ("max_speed", C.c_double), ("density_min", C.c_double), ("density_max", C.c_double),
        ("density_mean", C.c_double), ("pressure_min", C.c_double), ("pressure_max", C.c_double),
        ("pressure_mean", C.c_double), ("density_hist", C.c_int64 * DIAG_HIST_BINS),
        ("max_cell_count", C.c_int64), ("nonempty_cells", C.c_int64),
    ]


class SlabIpc(C.Structure):
    _fields_ = [("mem", (C.c_ubyte * 64) * 7), ("device", C.c_int32), ("ghost_capacity", C.c_int32)]


SLAB_INFO = ("n_owned", "n_first", "n_last", "n_ghost_below", "n_ghost_above", "errors",
             "migrants_in_below", "migrants_in_above")


class WcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"wc_sph error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libwc_sph.so (raises loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m watercube_b200.build` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.wc_abi_version.restype = C.c_int
        L.wc_last_error.restype = C.c_char_p
        sig = {
            "wc_default_params": [C.POINTER(Params)],
            "wc_default_step_params": [C.POINTER(StepParams)],
            "wc_derive": [C.POINTER(Params), C.POINTER(Derived)],
            "wc_device_count": [C.POINTER(C.c_int32)],
            "wc_create": [C.POINTER(Params), C.POINTER(vp)],
            "wc_destroy": [vp],
            "wc_get_derived": [vp, C.POINTER(Derived)],
            "wc_upload_particles": [vp, vp, i32],
            "wc_download_particles": [vp, i32, vp],
            "wc_step": [vp, f32, C.POINTER(StepParams)],
            "wc_step_host": [vp, f32, C.POINTER(StepParams), vp, C.c_int32, vp],
            "wc_step_export": [vp, f32, C.POINTER(StepParams), vp],
            "wc_sort_only": [vp],
            "wc_density_only": [vp, C.POINTER(StepParams)],
            "wc_update_only": [vp, f32, C.POINTER(StepParams)],
            "wc_advect_only": [vp, f32],
            "wc_download_cells": [vp, vp, vp, vp, vp, vp],
            "wc_download_forces": [vp, vp],
            "wc_upload_sorted": [vp, vp, i32],
            "wc_device_ptrs": [vp, C.POINTER(DeviceView)],
            "wc_export_aos_device": [vp, i32, vp],
            "wc_sync": [vp],
            "wc_stage_times": [vp, C.POINTER(C.c_float * NUM_STAGES)],
            "wc_launch_count": [vp, C.POINTER(C.c_uint64)],
            "wc_slab_get_view": [vp, C.POINTER(SlabView)],
            "wc_slab_clear_recv": [vp, i32],
            "wc_slab_sort_count": [vp],
            "wc_slab_sync_info": [vp, C.POINTER(C.c_int32 * 8)],
            "wc_slab_reorder": [vp],
            "wc_slab_density": [vp, C.POINTER(StepParams)],
            "wc_slab_update": [vp, f32, C.POINTER(StepParams)],
            "wc_slab_ipc_export": [vp, C.POINTER(SlabIpc)],
            "wc_slab_peer_open": [vp, i32, C.POINTER(SlabIpc)],
            "wc_slab_peer_attach": [vp, i32, vp],
            "wc_diagnose": [vp, i32, f32, C.POINTER(Diagnostics)],
            "wc_slab_step_peer": [vp, f32, C.POINTER(StepParams), C.POINTER(C.c_int32 * 8)],
            "wc_slab_step_peer_host": [vp, f32, C.POINTER(StepParams), vp, i32, vp, i32,
                                       C.POINTER(C.c_int32 * 8)],
            "wc_default_physics": [C.POINTER(Physics)],
            "wc_set_physics": [vp, C.POINTER(Physics)],
            "wc_get_physics": [vp, C.POINTER(Physics)],
            "wc_get_num_particles": [vp, C.POINTER(C.c_int32)],
        }
        for name, argtypes in sig.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = L
    return _lib


def check(rc):
    if rc != WC_OK:
        raise WcError(rc, lib().wc_last_error().decode(errors="replace"))


def default_params(**kw) -> Params:
    p = Params()
    check(lib().wc_default_params(C.byref(p)))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_step_params(**kw) -> StepParams:
    sp = StepParams()
    check(lib().wc_default_step_params(C.byref(sp)))
    for k, v in kw.items():
        if k in ("gravity", "mouse_origin", "mouse_dir"):
            getattr(sp, k)[:] = [float(x) for x in v]
        else:
            setattr(sp, k, v)
    return sp


def derive(p: Params) -> Derived:
    d = Derived()
    check(lib().wc_derive(C.byref(p), C.byref(d)))
    return d


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _as_aos(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype == PARTICLE_DTYPE:
        return a.view(np.float32).reshape(-1, 8)
    return np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 8)


class Fluid:
    """Python mirror of core::Fluid's setup/update/particle-buffer surface
    (src/core/Fluid.h:32-59) on top of the C-ABI."""

    def __init__(self, num_particles=80000, grid_res=21, size=1.0, particle_radius=0.01,
                 time_scale=0.012, device=0, flags=0, stream=None, capacity=0,
                 neighbour_list_words=0, slab=None, **step_kw):
        self.params = default_params(num_particles=num_particles, grid_res=grid_res, size=size,
                                     particle_radius=particle_radius, time_scale=time_scale,
                                     device=device, flags=flags, capacity=capacity,
                                     neighbour_list_words=neighbour_list_words, stream=stream)
        if slab is not None:  # (z_begin, z_end, ghost_capacity, migrant_capacity)
            (self.params.slab_z_begin, self.params.slab_z_end, self.params.slab_ghost_capacity,
             self.params.slab_migrant_capacity) = [int(x) for x in slab]
        self.step_params = default_step_params(**step_kw)
        self._h = C.c_void_p()
        check(lib().wc_create(C.byref(self.params), C.byref(self._h)))
        self.derived = Derived()
        check(lib().wc_get_derived(self._h, C.byref(self.derived)))

    # -- extended physics (wc_physics; not reference behaviour)
    def set_physics(self, flags, **values) -> Physics:
        """Defaults + overrides (surface_tension, surface_threshold, wall_stiffness,
        wall_distance, wall_rest_density); takes effect from the next step on."""
        ph = Physics()
        check(lib().wc_default_physics(C.byref(ph)))
        ph.flags = int(flags)
        for k, v in values.items():
            setattr(ph, k, v)
        check(lib().wc_set_physics(self._h, C.byref(ph)))
        return ph

    def physics(self) -> Physics:
        ph = Physics()
        check(lib().wc_get_physics(self._h, C.byref(ph)))
        return ph

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().wc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- particle-buffer surface
    @property
    def num_particles(self) -> int:
        n = C.c_int32()
        check(lib().wc_get_num_particles(self._h, C.byref(n)))
        return int(n.value)

    def upload(self, particles):
        """util::setParticles into buffer 1.  Accepts ndarray or a raw host pointer + n."""
        if isinstance(particles, tuple):
            ptr, n = particles
            check(lib().wc_upload_particles(self._h, C.c_void_p(ptr), int(n)))
            return
        a = _as_aos(particles)
        check(lib().wc_upload_particles(self._h, _ptr(a), a.shape[0]))

    def upload_sorted(self, particles):
        a = _as_aos(particles)
        check(lib().wc_upload_sorted(self._h, _ptr(a), a.shape[0]))

    def download(self, which=1, out=None) -> np.ndarray:
        """util::getParticles: which=1 current state, which=2 sorted with density/pressure."""
        if isinstance(out, tuple):
            check(lib().wc_download_particles(self._h, int(which), C.c_void_p(out[0])))
            return None
        n = self.num_particles
        if out is None:
            out = np.empty((n, 8), np.float32)
        check(lib().wc_download_particles(self._h, int(which), _ptr(out)))
        return out

    # -- stepping
    def step(self, frame_dt=1.0 / 60.0):
        """Fluid::update(double time)"""
        check(lib().wc_step(self._h, float(frame_dt), C.byref(self.step_params)))

    update = step

    def step_host(self, host_in, host_out, frame_dt=1.0 / 60.0):
        """setParticles -> Fluid::update -> getParticles as one call (wc_step_host).
        host_in: (pointer, n) or None (step the resident state); host_out: raw pointer to
        n * 32 bytes -- page-locked memory takes the fused device-to-host path."""
        ptr, n = host_in if host_in is not None else (None, 0)
        check(lib().wc_step_host(self._h, float(frame_dt), C.byref(self.step_params),
                                 C.c_void_p(ptr), int(n), C.c_void_p(host_out)))

    def step_export(self, aos_dst, frame_dt=1.0 / 60.0):
        """Fluid::update with the new buffer 1 also stored as 32-byte AoS records into the
        device-addressable buffer at raw pointer `aos_dst` (wc_step_export)."""
        check(lib().wc_step_export(self._h, float(frame_dt), C.byref(self.step_params),
                                   C.c_void_p(aos_dst)))

    def export_aos_device(self, which, device_ptr):
        """Pack buffer `which` (1 or 2) as 32-byte AoS into device memory (wc_export_aos_device)."""
        check(lib().wc_export_aos_device(self._h, int(which), C.c_void_p(device_ptr)))

    def sort_only(self):
        check(lib().wc_sort_only(self._h))

    def density_only(self):
        check(lib().wc_density_only(self._h, C.byref(self.step_params)))

    def update_only(self, frame_dt=1.0 / 60.0):
        check(lib().wc_update_only(self._h, float(frame_dt), C.byref(self.step_params)))

    def advect_only(self, frame_dt=1.0 / 60.0):
        """Fluid::runAdvectProg (dead in the reference, Fluid.cpp:351)"""
        check(lib().wc_advect_only(self._h, float(frame_dt)))

    def sync(self):
        check(lib().wc_sync(self._h))

    def diagnose(self, which=1) -> dict:
        """wc_diagnose: on-device invalid counts, conserved quantities, density histogram."""
        d = Diagnostics()
        check(lib().wc_diagnose(self._h, int(which), float(self.step_params.rest_density),
                                C.byref(d)))
        out = {}
        for name, ctype in Diagnostics._fields_:
            v = getattr(d, name)
            out[name] = np.array(v) if isinstance(v, C.Array) else v
        return out

    # -- z-slab mode (see include/wc_sph.h, wc_slab_*)
    def slab_view(self) -> SlabView:
        v = SlabView()
        check(lib().wc_slab_get_view(self._h, C.byref(v)))
        return v

    def slab_clear_recv(self, direction):
        check(lib().wc_slab_clear_recv(self._h, int(direction)))

    def slab_sort_count(self):
        check(lib().wc_slab_sort_count(self._h))

    def slab_sync_info(self) -> dict:
        info = (C.c_int32 * 8)()
        check(lib().wc_slab_sync_info(self._h, C.byref(info)))
        return dict(zip(SLAB_INFO, [int(x) for x in info]))

    def slab_reorder(self):
        check(lib().wc_slab_reorder(self._h))

    def slab_density(self):
        check(lib().wc_slab_density(self._h, C.byref(self.step_params)))

    def slab_update(self, frame_dt=1.0 / 60.0):
        check(lib().wc_slab_update(self._h, float(frame_dt), C.byref(self.step_params)))

    def slab_step_peer(self, frame_dt=1.0 / 60.0, wait=True):
        """The five slab phases as one call (all neighbours attached).  wait=True: returns the
        step's info; wait=False: only queues the step (no host wait anywhere in it)."""
        if not wait:
            check(lib().wc_slab_step_peer(self._h, float(frame_dt), C.byref(self.step_params), None))
            return None
        info = (C.c_int32 * 8)()
        check(lib().wc_slab_step_peer(self._h, float(frame_dt), C.byref(self.step_params),
                                      C.byref(info)))
        return dict(zip(SLAB_INFO, [int(x) for x in info]))

    def slab_step_peer_host(self, host_in, host_out, out_capacity, frame_dt=1.0 / 60.0) -> dict:
        """wc_slab_step_peer_host: host AoS in (pointer, n) or None, host AoS out (raw pointer)."""
        ptr, n = host_in if host_in is not None else (None, 0)
        info = (C.c_int32 * 8)()
        check(lib().wc_slab_step_peer_host(self._h, float(frame_dt), C.byref(self.step_params),
                                           C.c_void_p(ptr), int(n), C.c_void_p(host_out),
                                           int(out_capacity), C.byref(info)))
        return dict(zip(SLAB_INFO, [int(x) for x in info]))

    # -- peer-memory exchange (include/wc_sph.h, wc_slab_peer_*)
    def slab_ipc_export(self) -> bytes:
        blob = SlabIpc()
        check(lib().wc_slab_ipc_export(self._h, C.byref(blob)))
        return bytes(blob)

    def slab_peer_open(self, direction, blob: bytes):
        check(lib().wc_slab_peer_open(self._h, int(direction),
                                      C.byref(SlabIpc.from_buffer_copy(blob))))

    def slab_peer_attach(self, direction, other: "Fluid"):
        check(lib().wc_slab_peer_attach(self._h, int(direction), other._h))

    # -- inspection
    def cells(self, neighbour_counts=False):
        n, nb = self.num_particles, int(self.derived.num_bins)
        out = dict(cell_ids=np.empty(n, np.uint32), counts=np.empty(nb, np.uint32),
                   offsets=np.empty(nb, np.uint32), perm=np.empty(n, np.uint32))
        nc = np.empty(n, np.uint32) if neighbour_counts else None
        check(lib().wc_download_cells(self._h, _ptr(out["cell_ids"]), _ptr(out["counts"]),
                                      _ptr(out["offsets"]), _ptr(out["perm"]), _ptr(nc)))
        if neighbour_counts:
            out["neighbour_counts"] = nc
        return out

    def forces(self) -> np.ndarray:
        F = np.empty((self.num_particles, 3), np.float32)
        check(lib().wc_download_forces(self._h, _ptr(F)))
        return F

    def view(self) -> DeviceView:
        v = DeviceView()
        check(lib().wc_device_ptrs(self._h, C.byref(v)))
        return v

    def stage_times(self):
        ms = (C.c_float * NUM_STAGES)()
        check(lib().wc_stage_times(self._h, C.byref(ms)))
        return dict(zip(STAGES, [float(x) for x in ms]))

    def launch_count(self) -> int:
        v = C.c_uint64()
        check(lib().wc_launch_count(self._h, C.byref(v)))
        return int(v.value)
