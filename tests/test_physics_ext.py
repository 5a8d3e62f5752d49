"""Extended physics (what the reference's report lists as future work --
wall particles after Harada et al., surface tension after Yan et al. / Mueller's colour field).

There is no reference code for it, so there is nothing to be bit-compatible with: the oracle
(oracle/wc_oracle.h, WCO_PHYS_*) defines the arithmetic.  The CPU tests below pin that
definition to closed forms and physical properties; the GPU tests check the CUDA path
(include/wc_sph.h wc_set_physics) against the oracle with the tolerances of
tests/test_gpu_parity.py.  With flags == 0 everything is the reference's step, bit for bit.
"""
import numpy as np
import pytest

from watercube_b200 import scenes

f32 = np.float32
FRAME_DT = 1.0 / 60.0
WALL, TENSION = 1, 2


def scene_from_positions(pos, size, grid_res, radius=0.01, name="ext"):
    P = np.zeros((pos.shape[0], 8), f32)
    P[:, :3] = pos
    return scenes.Scene(name, P, float(size), int(grid_res), float(radius))


def oracle_stages(oracle, sc, **overrides):
    p = oracle.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res,
                              particle_radius=sc.particle_radius, **overrides)
    d = oracle.derive(p)
    s = oracle.sort(sc.particles, d.bin_size, p.grid_res)
    nt = oracle.host_threads()
    P, nc = oracle.density(s["sorted"], s["counts"], s["offsets"], p, nthreads=nt)
    dt = f32(FRAME_DT) * f32(p.time_scale)
    out, F = oracle.update(P, s["counts"], s["offsets"], p, dt, nthreads=nt)
    return dict(s, params=p, derived=d, sorted=oracle.as_f32(P), neighbour_counts=nc, force=F,
                out=oracle.as_f32(out), dt=float(dt))


# ------------------------------------------------------------------ the oracle's definition
def test_wall_weight_is_the_half_space_integral_of_the_density_kernel(oracle):
    """One particle at distance s from one wall: density = self term + wall_rest_density x the
    integral of W_poly6 over the half space behind the wall (quadrature here, closed form there):
    half of wall_rest_density on the wall, nothing from a distance h on."""
    size, G = 1.0, 21
    rho_w = 1234.0
    for frac in (0.0005, 0.1, 0.37, 0.5, 0.8, 0.999, 1.5):
        p0 = oracle.default_params(num_particles=1, size=size, grid_res=G)
        d = oracle.derive(p0)
        h = float(d.kernel_radius)
        s = frac * h
        sc = scene_from_positions(np.array([[s, 0.5, 0.5]], f32), size, G)
        o = oracle_stages(oracle, sc, physics_flags=WALL, wall_rest_density=rho_w)
        self_term = float(d.particle_mass) * float(d.poly6_const) * h ** 6
        # quadrature: slabs z in [s, h], disc of radius sqrt(h^2 - z^2): pi C (h^2 - z^2)^4 / 4
        z = np.linspace(min(float(f32(s)), h), h, 200001)
        want = rho_w * np.trapezoid(np.pi * float(d.poly6_const) * (h * h - z * z) ** 4 / 4.0, z)
        got = float(o["sorted"][0, 3]) - self_term
        assert got == pytest.approx(want, rel=2e-4, abs=2e-4 * rho_w), frac
        if frac < 0.001:
            assert got == pytest.approx(0.5 * rho_w, rel=2e-3)   # poly6 integrates to one
        if frac > 1.0:
            assert abs(got) < 1e-3            # (fp32 rounding of the self term)


def lattice_box(spacing, size):
    k = int(round(size / spacing))
    ax = (np.arange(k) + 0.5) * spacing
    return np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3).astype(f32), k


def test_wall_particles_give_a_lattice_the_inside_density_at_the_wall(oracle):
    """The point of Harada's wall particles: a resting lattice that fills the box up to the walls
    sees (nearly) the same density in its outermost layer as inside, because the wall stands in
    for the missing neighbours.  The reference's pseudo wall term does not do that."""
    size, G, spacing = 0.64, 13, 0.02
    pos, k = lattice_box(spacing, size)
    sc = scene_from_positions(pos, size, G)
    m = float(oracle.derive(oracle.default_params(num_particles=sc.n, size=size, grid_res=G)).particle_mass)
    o = oracle_stages(oracle, sc, physics_flags=WALL, wall_rest_density=m / spacing ** 3)
    ref = oracle_stages(oracle, sc)
    idx = np.rint(o["sorted"][:, :3] / spacing - 0.5).astype(int)
    inside = np.all((idx >= 3) & (idx <= k - 4), axis=1)
    face = (idx[:, 0] == 0) & np.all((idx[:, 1:] >= 3) & (idx[:, 1:] <= k - 4), axis=1)
    assert inside.sum() > 1000 and face.sum() > 100
    rho_in = float(o["sorted"][inside, 3].mean())
    assert float(np.abs(o["sorted"][face, 3] / rho_in - 1.0).max()) < 0.03
    assert float(np.abs(ref["sorted"][face, 3] / rho_in - 1.0).min()) > 0.2   # the quirk it replaces
    # the pressure never carries a wall term (density.comp:133, Q4) -- unchanged by the flag
    np.testing.assert_array_equal(o["sorted"][:, 7], ref["sorted"][:, 7])


def test_wall_push_undoes_the_given_fraction_of_the_penetration_in_one_step(oracle):
    size, G = 1.0, 21
    d_w, kappa = 0.01, 0.5
    pos = np.array([[0.004, 0.5, 0.5], [0.5, size - 0.007, 0.5], [0.5, 0.5, 0.02]], f32)
    sc = scene_from_positions(pos, size, G)
    o = oracle_stages(oracle, sc, physics_flags=WALL, wall_distance=d_w, wall_stiffness=kappa,
                      gravity=[0.0, 0.0, 0.0])
    order = np.argsort(o["perm"])                 # back to input order
    out = o["out"][order]
    assert out[0, 0] - pos[0, 0] == pytest.approx(kappa * (d_w - 0.004), rel=1e-4)
    assert pos[1, 1] - out[1, 1] == pytest.approx(kappa * (d_w - 0.007), rel=1e-3)
    np.testing.assert_array_equal(out[2, :3], pos[2])          # beyond the rest distance: no push
    np.testing.assert_array_equal(out[0, 1:3], pos[0, 1:3])    # only along the wall normal


def blob(radius, spacing, centre):
    k = int(np.ceil(radius / spacing)) + 1
    ax = np.arange(-k, k + 1) * spacing
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    g = g[np.linalg.norm(g, axis=1) <= radius]
    return (g + centre).astype(f32)


def test_surface_tension_pulls_the_surface_of_a_blob_inward_and_leaves_its_inside_alone(oracle):
    size, G, spacing, R = 1.0, 21, 0.02, 0.16
    centre = np.array([0.5, 0.5, 0.5])
    sc = scene_from_positions(blob(R, spacing, centre), size, G)
    kw = dict(gravity=[0.0, 0.0, 0.0])
    ref = oracle_stages(oracle, sc, **kw)
    o = oracle_stages(oracle, sc, physics_flags=TENSION, surface_tension=50.0,
                      surface_threshold=7.0, **kw)
    h = float(o["derived"].kernel_radius)
    np.testing.assert_array_equal(o["sorted"], ref["sorted"])   # the density pass is untouched
    dF = o["force"].astype(np.float64) - ref["force"].astype(np.float64)
    to_centre = centre - o["sorted"][:, :3].astype(np.float64)
    r = np.linalg.norm(to_centre, axis=1)
    inner = r < R - h - spacing
    assert inner.sum() > 50
    np.testing.assert_array_equal(o["force"][inner], ref["force"][inner])   # |n| = 0 by symmetry
    shell = r > R - 0.5 * spacing
    pulled = np.einsum("ij,ij->i", dF[shell], to_centre[shell] / r[shell, None])
    assert shell.sum() > 200 and np.all(pulled > 0.0)
    # along the normal: the tangential part of the pull is small
    assert np.median(pulled / np.linalg.norm(dF[shell], axis=1)) > 0.95
    assert np.abs(dF.sum(0)).max() < 1e-3 * np.abs(dF).sum(0).max()          # symmetric blob: no net force


def test_flags_zero_is_the_reference_step_whatever_the_other_values(oracle):
    sc = scenes.dam_break(3000, seed=3, size=0.3, grid_res=6)
    a = oracle_stages(oracle, sc)
    b = oracle_stages(oracle, sc, physics_flags=0, surface_tension=123.0, surface_threshold=0.5,
                      wall_stiffness=0.9, wall_distance=0.03, wall_rest_density=77.0)
    for key in ("sorted", "force", "out"):
        np.testing.assert_array_equal(a[key], b[key])


# ------------------------------------------------------------------ the CUDA path
RTOL_RHO, RTOL_P = 1e-5, 3e-5
C_EPS_F, EPS32 = 24.0, 2.0 ** -24
PHYS = dict(surface_tension=50.0, surface_threshold=7.0, wall_stiffness=0.5, wall_distance=0.01,
            wall_rest_density=9000.0)


@pytest.fixture(scope="module")
def capi():
    from watercube_b200 import capi as m

    m.lib()
    return m


def gpu_stages(capi, sc, flags, simple, **step_kw):
    fl_flags = capi.FLAG_DEBUG_OUTPUTS | (capi.FLAG_SIMPLE_KERNELS if simple else 0)
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size,
                    particle_radius=sc.particle_radius, flags=fl_flags, **step_kw) as fl:
        fl.set_physics(flags, **PHYS)
        assert fl.physics().flags == flags
        fl.upload(sc.particles)
        fl.sort_only()
        res = fl.cells()
        fl.density_only()
        res["neighbour_counts"] = fl.cells(neighbour_counts=True)["neighbour_counts"]
        res["sorted"] = fl.download(2)
        fl.update_only(FRAME_DT)
        res["force"] = fl.forces()
        res["out"] = fl.download(1)
    return res


def threshold_straddlers(oracle, sc, flags, ref):
    """Particles whose colour-field gradient is within 0.1 % of the surface threshold: the on/off
    decision may legitimately differ between two fp32 evaluations."""
    if not flags & TENSION:
        return np.zeros(sc.n, bool)
    lo = oracle_stages(oracle, sc, physics_flags=flags,
                       **dict(PHYS, surface_threshold=PHYS["surface_threshold"] * 0.999))
    hi = oracle_stages(oracle, sc, physics_flags=flags,
                       **dict(PHYS, surface_threshold=PHYS["surface_threshold"] * 1.001))
    return np.any(lo["force"] != hi["force"], axis=1)


@pytest.mark.gpu
@pytest.mark.parametrize("simple", [False, True], ids=["tiled", "simple"])
@pytest.mark.parametrize("flags", [WALL, TENSION, WALL | TENSION], ids=["wall", "tension", "both"])
@pytest.mark.parametrize("n", [20000, 200000])
def test_extended_physics_against_the_oracle(capi, oracle, n, flags, simple):
    sc = scenes.dam_break(n, seed=13)
    ref = oracle_stages(oracle, sc, physics_flags=flags, **PHYS)
    got = gpu_stages(capi, sc, flags, simple)
    for key in ("cell_ids", "counts", "offsets", "perm", "neighbour_counts"):
        np.testing.assert_array_equal(got[key], ref[key], err_msg=key)
    r_rho, r_p = ref["sorted"][:, 3], ref["sorted"][:, 7]
    if flags & WALL:   # the wall weight is positive: no cancellation to allow for
        assert np.all(np.abs(got["sorted"][:, 3] - r_rho) <= RTOL_RHO * np.abs(r_rho))
    assert np.all(np.abs(got["sorted"][:, 7] - r_p) <= RTOL_P * (np.abs(r_p) + 100.0))
    # forces: a few fp32 units of the fp64 oracle's term scale, component by component
    P = oracle.as_particles(ref["sorted"])
    _, _, _, S = oracle.update_f64_scaled(P, P["density"].astype(np.float64),
                                          P["pressure"].astype(np.float64), ref["counts"],
                                          ref["offsets"], ref["params"], f32(ref["dt"]),
                                          nthreads=oracle.host_threads())
    keep = ~threshold_straddlers(oracle, sc, flags, ref)
    assert keep.mean() > 0.99
    units = np.abs(got["force"][keep] - ref["force"][keep]) / (EPS32 * np.maximum(S[keep], 1e-300))
    assert units.max() <= C_EPS_F, float(units.max())
    # the extended terms are really in there (not the reference's step by accident)
    plain = oracle_stages(oracle, sc)
    assert np.abs(ref["force"] - plain["force"]).max() > 1e-3 * np.abs(plain["force"]).max()
    vmax = max(np.abs(ref["out"][:, 4:7]).max(), 1e-3)
    assert np.abs(got["out"][keep, 4:7] - ref["out"][keep, 4:7]).max() <= 2e-5 * max(vmax, 50.0)


@pytest.mark.gpu
@pytest.mark.parametrize("simple", [False, True], ids=["tiled", "simple"])
def test_extended_physics_against_the_golden_fixture(capi, simple):
    """tests/golden/dam_break_4096_ext.npz (frozen oracle output; does not need the oracle build):
    integers bit-exact, density / pressure / force / state within the fp32 bounds of
    tests/test_gpu_parity.py, for the particles the fixture marks as clear of the surface
    threshold."""
    import os

    from tests.golden import make_golden

    name = "dam_break_4096_ext"
    factory, overrides, _ = make_golden.CASES[name]
    sc = factory()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    phys = {k: v for k, v in overrides.items() if k != "physics_flags"}
    assert phys == PHYS
    got = gpu_stages(capi, sc, overrides["physics_flags"], simple)
    for key in ("cell_ids", "counts", "offsets", "perm", "neighbour_counts"):
        np.testing.assert_array_equal(got[key], g[key], err_msg=key)
    assert np.all(np.abs(got["sorted"][:, 3] - g["density"]) <= RTOL_RHO * np.abs(g["density"]))
    assert np.all(np.abs(got["sorted"][:, 7] - g["pressure"]) <= RTOL_P * (np.abs(g["pressure"]) + 100.0))
    keep = g["threshold_clear"]
    assert keep.mean() > 0.99
    F = g["force"][keep]
    fn = np.abs(F).max(1)
    err = np.abs(got["force"][keep] - F).max(1)
    assert np.all(err <= 1e-5 * fn + 1e-6 * fn.max())                 # SURVEY 7.4's per-particle bound
    assert np.abs(got["out"][keep, 4:7] - g["out"][keep, 4:7]).max() <= 2e-5 * 50.0
    assert np.abs(got["out"][keep, :3] - g["out"][keep, :3]).max() <= 2e-6


@pytest.mark.gpu
def test_flags_zero_on_the_gpu_is_the_reference_step(capi):
    sc = scenes.dam_break(50000, seed=2)
    outs = []
    for touch in (False, True):
        with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size,
                        particle_radius=sc.particle_radius) as fl:
            if touch:
                fl.set_physics(0, surface_tension=99.0, wall_distance=0.05)
            fl.upload(sc.particles)
            for _ in range(3):
                fl.step(FRAME_DT)
            outs.append(fl.download(1))
    np.testing.assert_array_equal(outs[0], outs[1])


@pytest.mark.gpu
def test_unknown_flag_bits_and_bad_values_are_refused(capi):
    with capi.Fluid(num_particles=10, grid_res=4, size=0.2) as fl:
        with pytest.raises(capi.WcError):
            fl.set_physics(4)
        with pytest.raises(capi.WcError):
            fl.set_physics(WALL, wall_distance=float("nan"))
        with pytest.raises(capi.WcError):
            fl.set_physics(WALL, wall_stiffness=-1.0)
        assert fl.physics().flags == 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_extended_physics_z_slabs_equal_the_whole_grid(capi, world):
    """The walls are the global box's and the colour field reads the ghosts' densities: slabs
    with the flags set reproduce the whole-grid run with the flags set, bit for bit."""
    from watercube_b200 import slab

    sc = scenes.dam_break(80000, seed=11)
    kw = dict(grid_res=sc.grid_res, size=sc.size, particle_radius=sc.particle_radius)
    steps = 6
    with capi.Fluid(num_particles=sc.n, **kw) as fl:
        fl.set_physics(WALL | TENSION, **PHYS)
        fl.upload(sc.particles)
        for _ in range(steps):
            fl.step(FRAME_DT)
        want1, want2 = fl.download(1), fl.download(2)
    d = capi.derive(capi.default_params(num_particles=sc.n, **kw))
    hist = np.bincount(slab.layer_of(sc.particles[:, 2], d.bin_size, sc.grid_res),
                       minlength=sc.grid_res)
    cuts = slab.slab_cuts(hist, world)
    parts = slab.decompose(sc.particles, cuts, d.bin_size, sc.grid_res)
    backends = []
    for r in range(world):
        b = slab.CudaSlabBackend(kw, cuts[r], cuts[r + 1], capacity=sc.n, ghost_capacity=sc.n,
                                 migrant_capacity=8192)
        b.fluid.set_physics(WALL | TENSION, **PHYS)
        b.upload(parts[r])
        backends.append(b)
    slab.attach_peers_local(backends)
    slab.run_steps_peer_async(backends, FRAME_DT, steps=steps)
    np.testing.assert_array_equal(np.concatenate([b.download(2) for b in backends]), want2)
    np.testing.assert_array_equal(np.concatenate([b.download(1) for b in backends]), want1)
    for b in backends:
        b.close()


@pytest.mark.gpu
def test_extended_physics_long_run_stays_bounded_and_off_the_walls(capi):
    sc = scenes.dam_break(100000, seed=5)
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size,
                    particle_radius=sc.particle_radius) as fl:
        fl.set_physics(WALL | TENSION)
        fl.upload(sc.particles)
        for _ in range(300):
            fl.step(FRAME_DT)
        out = fl.download(1)
    assert np.isfinite(out).all()
    pos, vel = out[:, :3], out[:, 4:7]
    assert pos.min() >= 0.001 and pos.max() <= sc.size - 0.001
    assert np.abs(vel).max() <= 50.0
    # the push keeps the fluid off the walls: hardly anything sits on the clamp border
    assert np.mean(np.minimum(pos, sc.size - pos).min(1) < 0.002) < 1e-3
