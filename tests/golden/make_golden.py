"""Generates tests/golden/*.npz from the CPU oracle:  python -m tests.golden.make_golden

The reference has no golden vectors of its own (parity unpinned) and
cannot run here, so these fixtures freeze the ORACLE's outputs (which tests/test_oracle.py
pins against closed forms and an independent numpy restatement).  They guard the oracle
against regressions and give the GPU parity tests a fixture that does not depend on the
oracle build on the GPU box.  Inputs are regenerated from watercube_b200.scenes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402
from watercube_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FRAME_DT = 1.0 / 60.0

EXT = {"physics_flags": 3, "surface_tension": 50.0, "surface_threshold": 7.0,
       "wall_stiffness": 0.5, "wall_distance": 0.01, "wall_rest_density": 9000.0}

CASES = {
    # name: (scene factory, param overrides, subsample stride for per-particle floats)
    "dam_break_4096": (lambda: scenes.dam_break(4096, seed=1, size=0.36, grid_res=7), {}, 1),
    "uniform_3000": (lambda: scenes.uniform_box(3000, size=0.4, h=0.04, seed=2),
                     {"gravity": [30.0, -900.0, 10.0], "rest_pressure": 25.0}, 1),
    "default_80000": (lambda: scenes.dam_break(80000, seed=0), {}, 97),
    # the extended physics (oracle WCO_PHYS_*: wall particles + surface tension).  The fixture
    # also carries `threshold_clear`: the particles whose colour-field gradient is not within
    # 0.1 % of the surface threshold (for the others the on/off decision may differ between
    # two fp32 evaluation orders).
    "dam_break_4096_ext": (lambda: scenes.dam_break(4096, seed=1, size=0.36, grid_res=7), EXT, 1),
}


def case_inputs(name):
    factory, overrides, stride = CASES[name]
    sc = factory()
    p = ob.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res,
                          particle_radius=sc.particle_radius, **overrides)
    return sc, p, stride


def run_case(name):
    sc, p, stride = case_inputs(name)
    d = ob.derive(p)
    s = ob.sort(sc.particles, d.bin_size, p.grid_res)
    P, nc = ob.density(s["sorted"], s["counts"], s["offsets"], p, nthreads=ob.max_threads())
    dt = np.float32(FRAME_DT) * np.float32(p.time_scale)
    out, F = ob.update(P, s["counts"], s["offsets"], p, dt, nthreads=ob.max_threads())
    return dict(
        cell_ids=s["cell_ids"], counts=s["counts"], offsets=s["offsets"], perm=s["perm"],
        neighbour_counts=nc,
        density=P["density"][::stride].copy(), pressure=P["pressure"][::stride].copy(),
        force=F[::stride].copy(), out=ob.as_f32(out)[::stride].copy(),
    )


def threshold_clear_mask(name):
    """Particles whose force does not change when the surface threshold moves by +-0.1 %: their
    surface on/off decision is the same in every fp32 evaluation order (all True without
    surface tension)."""
    factory, overrides, stride = CASES[name]
    if not overrides.get("physics_flags", 0) & 2:
        return None
    forces = []
    for scale in (0.999, 1.001):
        CASES["_probe"] = (factory, dict(overrides, surface_threshold=overrides["surface_threshold"] * scale), stride)
        forces.append(run_case("_probe")["force"])
        del CASES["_probe"]
    return np.all(forces[0] == forces[1], axis=1)


def main():
    for name in CASES:
        res = run_case(name)
        keep = threshold_clear_mask(name)
        if keep is not None:
            assert keep.mean() > 0.99
            res["threshold_clear"] = keep
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **res)
        print(name, {k: v.shape for k, v in res.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
