"""wc_diagnose: the on-device counterpart of the reference's visual / console inspection
(particle.vert:35-55 render modes, Sort::printGrids Sort.cpp:237-249, util::printParticles
util.cpp:113-126).  The checker is plain numpy over the downloaded buffers."""
import ctypes as C

import numpy as np
import pytest

from watercube_b200 import scenes

f32 = np.float32
FRAME_DT = 1.0 / 60.0


@pytest.fixture(scope="module")
def capi():
    from watercube_b200 import capi as m

    m.lib()
    return m


def numpy_diagnostics(P, size, mass, rho0, capi):
    """The same definitions as include/wc_sph.h wc_diagnostics, in fp64 numpy."""
    P = np.asarray(P, f32)
    x, rho, v, pres = P[:, 0:3], P[:, 3], P[:, 4:7], P[:, 7]
    with np.errstate(invalid="ignore", over="ignore"):
        inside = np.all((x >= 0) & (x <= f32(size)), axis=1)
        v2 = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]).astype(f32)
        finite = np.isfinite(v2) & np.isfinite(rho) & np.isfinite(pres)
        valid = inside & finite & (rho > 0)
    V, X = v[valid].astype(np.float64), x[valid].astype(np.float64)
    nv = int(valid.sum())
    ratio = (rho[valid] * f32(f32(1.0) / f32(rho0)) * f32(capi.DIAG_HIST_PER_UNIT)).astype(f32)
    bins = np.minimum(ratio.astype(np.int64), capi.DIAG_HIST_BINS - 1)
    bins[ratio >= capi.DIAG_HIST_BINS - 1] = capi.DIAG_HIST_BINS - 1
    return dict(
        particles=len(P), invalid=int((~valid).sum()), out_of_box=int((~inside).sum()),
        at_speed_clamp=int((np.abs(v[valid]) >= 50.0).any(axis=1).sum()),
        mass=mass * nv, momentum=mass * V.sum(axis=0), kinetic_energy=0.5 * mass * (V * V).sum(),
        centre_of_mass=X.mean(axis=0) if nv else np.zeros(3),
        max_speed=float(np.sqrt(v2[valid].max())) if nv else 0.0,
        density_min=float(rho[valid].min()) if nv else 0.0,
        density_max=float(rho[valid].max()) if nv else 0.0,
        density_mean=float(rho[valid].astype(np.float64).mean()) if nv else 0.0,
        pressure_min=float(pres[valid].min()) if nv else 0.0,
        pressure_max=float(pres[valid].max()) if nv else 0.0,
        pressure_mean=float(pres[valid].astype(np.float64).mean()) if nv else 0.0,
        density_hist=np.bincount(bins, minlength=capi.DIAG_HIST_BINS),
    )


def assert_same(got, ref):
    for key in ("particles", "invalid", "out_of_box", "at_speed_clamp"):
        assert got[key] == ref[key], key
    np.testing.assert_array_equal(got["density_hist"], ref["density_hist"])
    for key in ("density_min", "density_max", "pressure_min", "pressure_max"):
        assert got[key] == ref[key], key                      # exact: min / max of fp32 values
    for key in ("mass", "kinetic_energy", "density_mean", "pressure_mean", "max_speed"):
        assert got[key] == pytest.approx(ref[key], rel=1e-6, abs=1e-12), key   # max_speed: fp32 sqrt arg
    scale = max(float(np.abs(ref["momentum"]).max()), ref["kinetic_energy"] ** 0.5, 1e-12)
    np.testing.assert_allclose(got["momentum"], ref["momentum"], rtol=0, atol=1e-9 * scale + 1e-12)
    np.testing.assert_allclose(got["centre_of_mass"], ref["centre_of_mass"], rtol=1e-12, atol=1e-12)


def test_diagnose_rejects_bad_arguments(capi):
    d = capi.Diagnostics()
    assert capi.lib().wc_diagnose(None, 1, 500.0, C.byref(d)) == capi.WC_ERR_INVALID
    assert C.sizeof(capi.Diagnostics) == 8 * 4 + 8 * (1 + 3 + 1 + 3 + 1 + 6) + 8 * 32 + 16


@pytest.mark.gpu
def test_diagnose_matches_numpy_after_steps(capi):
    sc = scenes.dam_break(200_000, seed=4)
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size) as fl:
        fl.upload(sc.particles)
        for _ in range(12):
            fl.step(FRAME_DT)
        mass = float(fl.derived.particle_mass)
        for which in (1, 2):
            got = fl.diagnose(which)
            ref = numpy_diagnostics(fl.download(which), sc.size, mass, 500.0, capi)
            assert_same(got, ref)
            assert got["invalid"] == 0
            again = fl.diagnose(which)                         # fixed fold order: same bits
            for key in ("kinetic_energy", "density_mean", "pressure_mean"):
                assert again[key] == got[key]
            np.testing.assert_array_equal(again["momentum"], got["momentum"])
        cells = fl.cells()
        assert got["max_cell_count"] == int(cells["counts"].max())
        assert got["nonempty_cells"] == int((cells["counts"] > 0).sum())


@pytest.mark.gpu
def test_diagnose_counts_invalid_particles(capi):
    """What render modes 1 / 2 would paint red: outside the box, non-finite, density <= 0."""
    sc = scenes.dam_break(5000, seed=1, size=0.4, grid_res=8)
    P = sc.particles.copy()
    P[:, 0:3] = np.clip(P[:, 0:3], 0.001, sc.size - 0.001)   # the lattice's jitter dips below 0 (Q11)
    P[:, 3] = 750.0
    P[:, 7] = 10.0
    P[0, 0] = -0.01            # outside
    P[1, 2] = 0.41             # outside
    P[2, 1] = np.nan           # NaN position (counts as outside as well)
    P[3, 4] = np.inf           # non-finite velocity
    P[4, 3] = 0.0              # density not positive
    P[5, 3] = -3.0
    P[6, 4:7] = [50.0, 0.0, -50.0]   # at the clamp, valid
    P[7, 3] = 500.0 * 9.0      # beyond the histogram: last bin
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size) as fl:
        fl.upload(P)
        got = fl.diagnose(1)
        assert got["max_cell_count"] == -1 and got["nonempty_cells"] == -1   # nothing sorted yet
        ref = numpy_diagnostics(P, sc.size, float(fl.derived.particle_mass), 500.0, capi)
        assert_same(got, ref)
        assert (got["invalid"], got["out_of_box"], got["at_speed_clamp"]) == (6, 3, 1)
        assert got["density_hist"][6] == sc.n - 7 and got["density_hist"][-1] == 1
    with capi.Fluid(num_particles=0, capacity=16, grid_res=4, size=0.2) as fl:
        got = fl.diagnose(1)                                   # empty: zeros, not an error
        assert got["particles"] == 0 and got["mass"] == 0.0 and got["density_hist"].sum() == 0
