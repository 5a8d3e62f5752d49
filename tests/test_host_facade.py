"""The C++ host facade (core::Fluid / core::Sort / core::util over the C-ABI) and the
headless driver.  CPU part: it builds, exports the reference's class surface, generates the
same scene as scenes.dam_break bit for bit, and fails loudly without a device.  GPU part: the
frame loop through core::Fluid::update equals the same steps through the Python binding."""
import json
import os
import subprocess

import numpy as np
import pytest

from watercube_b200 import scenes

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "watercube_b200",
                    "host")


@pytest.fixture(scope="module")
def exe():
    from watercube_b200.host import build as hb

    return hb.build()


def test_facade_exports_reference_surface(exe):
    out = subprocess.run(["nm", "-DC", os.path.join(HOST, "libwc_core.so")], capture_output=True,
                         text=True, check=True).stdout
    for sym in ("core::Fluid::setup()", "core::Fluid::update(double)", "core::Fluid::numParticles(int)",
                "core::Fluid::gridRes(int)", "core::Fluid::size(float)",
                "core::Fluid::particleRadius(float)", "core::Fluid::viscosityCoefficient(float)",
                "core::Fluid::stiffness(float)", "core::Fluid::restDensity(float)",
                "core::Fluid::restPressure(float)", "core::Fluid::gravityStrength(float)",
                "core::Fluid::reset()", "core::Sort::run(core::Buffer, core::Buffer)",
                "core::Sort::prepareBuffers()", "core::Sort::numItems(int)", "core::Sort::binSize(float)",
                "core::util::getParticles(core::Buffer, int)", "core::util::getUints(core::Buffer, int)",
                "core::util::setParticles(core::Buffer"):
        assert sym in out, sym
    # the facade talks to the device through the C-ABI only: no CUDA runtime symbols of its own
    undefined = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(HOST, "libwc_core.so")],
                               capture_output=True, text=True, check=True).stdout
    assert "wc_step" in undefined and "wc_create" in undefined
    assert "cuda" not in undefined.lower()


@pytest.mark.parametrize("n", [1000, 80000])
def test_cpp_scene_generator_equals_python(exe, n, tmp_path):
    dump = tmp_path / "init.bin"
    subprocess.run([exe, "--particles", str(n), "--initial-only", "--dump", str(dump)], check=True)
    got = np.fromfile(dump, np.float32).reshape(-1, 8)
    assert np.array_equal(got, scenes.dam_break(n, seed=0).particles)


def test_headless_fails_loudly_without_device(exe):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    res = subprocess.run([exe, "--particles", "1000", "--steps", "1"], capture_output=True, text=True)
    assert res.returncode == 2
    assert "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_headless_frame_loop_equals_python_binding(exe, tmp_path):
    from watercube_b200 import capi

    steps = 20
    dump = tmp_path / "state.bin"
    res = subprocess.run([exe, "--steps", str(steps), "--dump", str(dump)], capture_output=True,
                         text=True)
    assert res.returncode == 0, res.stderr
    stats = json.loads(res.stdout.strip().splitlines()[-1])
    assert stats["particles"] == 80000 and stats["invalid"] == 0
    got = np.fromfile(dump, np.float32).reshape(-1, 8)
    sc = scenes.dam_break(80000, seed=0)
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size) as fl:
        fl.upload(sc.particles)
        for _ in range(steps):
            fl.step(1.0 / 60.0)
        ref = fl.download(1)
    assert np.array_equal(got, ref)   # same library, same inputs: bit-identical
    ke = 0.5 * 0.08 * float((ref[:, 4:7].astype(np.float64) ** 2).sum())
    assert abs(stats["kinetic_energy"] - ke) <= 1e-6 * max(ke, 1.0)
