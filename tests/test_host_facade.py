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


# ---------------------------------------------------------------- scene layer + checkpoints
SCENE_PROBE = r"""
#include <cstdio>
#include "core/Scene.h"
#include "core/util.h"
using namespace core;
struct Probe : BaseObject {
    int updates = 0, draws = 0, resets = 0; double last = 0;
    explicit Probe(const std::string& n) : BaseObject(n) {}
    void update(double t) override { updates++; last = t; }
    void draw() override { draws++; }
    void reset() override { resets++; }
};
int main(int argc, char** argv) {
    SceneRef scene = Scene::create();
    auto a = std::make_shared<Probe>("a"), b = std::make_shared<Probe>("b"), dup = std::make_shared<Probe>("a");
    bool ok = scene->addObject(a) && scene->addObject(b, false) && !scene->addObject(dup) &&
              !scene->addObject(BaseObjectRef());
    scene->update(0.25); scene->update(0.5); scene->draw(); scene->reset();
    ok = ok && scene->numObjects() == 2 && scene->exists("b") && !scene->exists("c") &&
         scene->getObject("a") == a && !scene->getObject("zz") && scene->getObjectFromIndex(1) == b &&
         !scene->getObjectFromIndex(2) && a->updates == 2 && b->updates == 2 && b->last == 0.5 &&
         a->draws == 1 && b->draws == 0 && a->resets == 1 && b->resets == 1;
    scene->clear();
    ok = ok && scene->numObjects() == 0 && !scene->exists("a");
    // checkpoint file round trip (host only)
    std::vector<Particle> ps(3);
    ps[1].position = vec3(1, 2, 3); ps[2].pressure = 7.5f;
    util::CheckpointHeader h = util::CheckpointHeader();
    h.grid_res = 21; h.size = 1.5f; h.particle_radius = 0.01f; h.time_scale = 0.012f; h.steps = 42; h.time = 0.7;
    h.stiffness = 60.0f; h.gravity_direction[1] = -1.0f; h.has_mouse_ray = 1; h.mouse_dir[2] = 1.0f;
    util::saveCheckpoint(argv[1], h, ps);
    util::CheckpointHeader g;
    std::vector<Particle> back = util::loadCheckpoint(argv[1], &g);
    ok = ok && back.size() == 3 && back[1].position.z == 3 && back[2].pressure == 7.5f && g.steps == 42 &&
         g.num_particles == 3 && g.size == 1.5f && g.time == 0.7 && g.version == 2 &&
         g.stiffness == 60.0f && g.gravity_direction[1] == -1.0f && g.has_mouse_ray == 1 &&
         g.mouse_dir[2] == 1.0f;
    bool threw = false;
    try { util::loadCheckpoint(argv[2], nullptr); } catch (const core::Error&) { threw = true; }
    if (argc > 3) {  // a version-1 file: 64-byte header, no step parameters
        util::CheckpointHeader v1;
        std::vector<Particle> old = util::loadCheckpoint(argv[3], &v1);
        ok = ok && v1.version == 1 && old.size() == 3 && old[2].pressure == 7.5f && v1.steps == 42;
    }
    std::printf(ok && threw ? "ok\n" : "FAILED\n");
    return ok && threw ? 0 : 1;
}
"""


def test_scene_registry_and_checkpoint_file(exe, tmp_path):
    """Scene: registration order, unique names, visible-only draw (src/core/Scene.cpp:20-68);
    checkpoint: header + AoS round trip, and a foreign file is rejected."""
    src = tmp_path / "probe.cpp"
    src.write_text(SCENE_PROBE)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"not a checkpoint" * 8)
    probe = tmp_path / "probe"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", HOST, str(src), "-o", str(probe),
                    "-L", HOST, "-lwc_core", f"-Wl,-rpath,{HOST}",
                    "-L", os.path.join(HOST, "..", "csrc"), "-lwc_sph",
                    f"-Wl,-rpath,{os.path.join(HOST, '..', 'csrc')}"], check=True)
    ckpt = tmp_path / "c.wcb"
    res = subprocess.run([str(probe), str(ckpt), str(bad)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = ckpt.read_bytes()
    assert raw[:6] == b"WCB200" and len(raw) == 160 + 3 * 32        # version-2 header
    got = np.frombuffer(raw, np.float32, offset=160).reshape(-1, 8)
    assert got[1, 2] == 3.0 and got[2, 7] == 7.5
    # a header that promises more particles than the file holds is rejected BEFORE any
    # allocation is sized from it (and a version-1 file, 64-byte header, still loads)
    lying = bytearray(raw)
    lying[12:16] = (2_000_000_000).to_bytes(4, "little")
    (tmp_path / "lying.wcb").write_bytes(bytes(lying))
    res = subprocess.run([str(probe), str(tmp_path / "c2.wcb"), str(tmp_path / "lying.wcb")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    v1 = bytearray(raw[:64] + raw[160:])
    v1[8:12] = (1).to_bytes(4, "little")
    (tmp_path / "v1.wcb").write_bytes(bytes(v1))
    res = subprocess.run([str(probe), str(tmp_path / "c3.wcb"), str(bad), str(tmp_path / "v1.wcb")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr


@pytest.mark.gpu
def test_checkpoint_restore_continues_bit_identically(exe, tmp_path):
    """10 steps + checkpoint + restore + 10 steps == 20 steps straight (buffer 1 keeps its order,
    and the stable sort only sees positions and order)."""
    straight, first, second = tmp_path / "s.bin", tmp_path / "a.wcb", tmp_path / "b.bin"
    subprocess.run([exe, "--particles", "30000", "--steps", "20", "--dump", str(straight)], check=True,
                   capture_output=True)
    subprocess.run([exe, "--particles", "30000", "--steps", "10", "--checkpoint", str(first)], check=True,
                   capture_output=True)
    res = subprocess.run([exe, "--restore", str(first), "--steps", "10", "--dump", str(second)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    stats = json.loads(res.stdout.strip().splitlines()[-1])
    assert stats["particles"] == 30000 and stats["total_steps"] == 20
    assert np.array_equal(np.fromfile(second, np.float32), np.fromfile(straight, np.float32))
    # the on-device reduction agrees with the host loop over the downloaded buffer
    assert stats["device"]["invalid"] == stats["invalid"] == 0
    assert stats["device"]["kinetic_energy"] == pytest.approx(stats["kinetic_energy"], rel=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 3])
def test_headless_decomposed_run_equals_single_device(exe, tmp_path, gpus):
    """North star: the host code of the decomposed scenes is C++ with the Fluid surface.
    `wc_headless --gpus N` = core::Fluid::devices(N): N z-slabs driven by ONE process and thread
    (wc_slab_peer_attach + asynchronous wc_slab_step_peer), buffer 1 read back as the slabs in z
    order -- bit-identical to the single-device run.  (On a box with fewer devices the slabs
    share them; the multi-GPU bench covers distinct devices.)"""
    one, many = tmp_path / "one.bin", tmp_path / "many.bin"
    args = ["--particles", "60000", "--size", "0.9", "--grid", "18", "--steps", "12"]
    subprocess.run([exe, *args, "--dump", str(one)], check=True, capture_output=True)
    res = subprocess.run([exe, *args, "--gpus", str(gpus), "--dump", str(many)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    stats = json.loads(res.stdout.strip().splitlines()[-1])
    assert stats["particles"] == 60000 and stats["invalid"] == 0
    assert stats["device"]["kinetic_energy"] == pytest.approx(stats["kinetic_energy"], rel=1e-9)
    assert np.array_equal(np.fromfile(many, np.float32), np.fromfile(one, np.float32))


@pytest.mark.gpu
def test_checkpoint_carries_the_step_parameters(exe, tmp_path):
    """Version-2 checkpoints store the per-step parameters: a run with non-default stiffness /
    viscosity / rest density restores and continues bit-identically WITHOUT the flags repeated
    (also across the decomposed driver: saved from 2 slabs, restored on one device)."""
    straight, first, second = tmp_path / "s.bin", tmp_path / "a.wcb", tmp_path / "b.bin"
    phys = ["--stiffness", "60", "--viscosity", "120", "--rest-density", "650"]
    subprocess.run([exe, "--particles", "30000", "--steps", "16", *phys, "--dump", str(straight)],
                   check=True, capture_output=True)
    subprocess.run([exe, "--particles", "30000", "--steps", "8", *phys, "--gpus", "2",
                    "--checkpoint", str(first)], check=True, capture_output=True)
    res = subprocess.run([exe, "--restore", str(first), "--steps", "8", "--dump", str(second)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert json.loads(res.stdout.strip().splitlines()[-1])["total_steps"] == 16
    assert np.array_equal(np.fromfile(second, np.float32), np.fromfile(straight, np.float32))


# ---------------------------------------------------------------- extended physics (wc_physics)
@pytest.mark.gpu
def test_headless_extended_physics_equals_binding_decomposed_and_restored(exe, tmp_path):
    """`--wall-particles --surface-tension S` = Fluid::wallParticles / surfaceTension: the same
    bits as the ctypes binding with wc_set_physics, the same again from two z-slabs, and a
    checkpoint carries the physics record (restored run continues without the flags repeated)."""
    from watercube_b200 import capi

    steps = 16
    phys = ["--wall-particles", "--wall-density", "9000", "--surface-tension", "40"]
    straight, slabs, ck, resumed = (tmp_path / n for n in ("s.bin", "m.bin", "a.wcb", "r.bin"))
    res = subprocess.run([exe, "--steps", str(steps), *phys, "--dump", str(straight)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = np.fromfile(straight, np.float32).reshape(-1, 8)
    sc = scenes.dam_break(80000, seed=0)
    with capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size) as fl:
        fl.set_physics(capi.PHYS_WALL_PARTICLES | capi.PHYS_SURFACE_TENSION, surface_tension=40.0,
                       wall_rest_density=9000.0)
        fl.upload(sc.particles)
        for _ in range(steps):
            fl.step(1.0 / 60.0)
        ref = fl.download(1)
        plain = capi.Fluid(num_particles=sc.n, grid_res=sc.grid_res, size=sc.size)
        plain.upload(sc.particles)
        for _ in range(steps):
            plain.step(1.0 / 60.0)
        assert not np.array_equal(plain.download(1), ref)      # the flags do change the run
        plain.close()
    assert np.array_equal(got, ref)
    subprocess.run([exe, "--steps", str(steps), *phys, "--gpus", "2", "--dump", str(slabs)],
                   check=True, capture_output=True)
    assert np.array_equal(np.fromfile(slabs, np.float32).reshape(-1, 8), ref)
    subprocess.run([exe, "--steps", str(steps // 2), *phys, "--checkpoint", str(ck)], check=True,
                   capture_output=True)
    subprocess.run([exe, "--restore", str(ck), "--steps", str(steps // 2), "--dump", str(resumed)],
                   check=True, capture_output=True)
    assert np.array_equal(np.fromfile(resumed, np.float32).reshape(-1, 8), ref)
