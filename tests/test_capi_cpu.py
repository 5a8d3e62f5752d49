"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/wc_sph.h
declares, its device-free entry points agree with the oracle, and -- having no CPU
fallback -- every compute entry point fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from watercube_b200 import build, capi as m

    build.build()
    m.lib()
    return m


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wc_sph.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    syms = declared_symbols()
    assert len(syms) >= 20
    L = C.CDLL(capi.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in wc_sph.h but not exported"
    assert sorted(capi.EXPORTS) == syms
    assert L.wc_abi_version() == 1


def test_header_is_plain_c_and_sizes_agree(capi, tmp_path):
    """The drop-in boundary is a C ABI: include/wc_sph.h compiles as strict C99, and the struct
    sizes the C compiler sees are the ones the ctypes mirror uses."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "wc_sph.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(wc_params), '
        'sizeof(wc_step_params), sizeof(wc_derived), sizeof(wc_particle), sizeof(wc_device_view), '
        'sizeof(wc_slab_view), sizeof(wc_slab_ipc), sizeof(wc_diagnostics)); return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                    "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True,
                                          check=True).stdout.split()]
    want = [C.sizeof(t) for t in (capi.Params, capi.StepParams, capi.Derived)] + \
           [capi.PARTICLE_DTYPE.itemsize] + \
           [C.sizeof(t) for t in (capi.DeviceView, capi.SlabView, capi.SlabIpc, capi.Diagnostics)]
    assert got == want


def test_struct_layouts_match_header(capi):
    assert C.sizeof(capi.Params) == 64          # 13 x 4 bytes, padding, pointer
    assert C.sizeof(capi.StepParams) == 52
    assert C.sizeof(capi.Derived) == 32
    assert capi.PARTICLE_DTYPE.itemsize == 32   # src/core/util.h:29-35


def test_defaults_match_reference_constructor(capi):
    p, sp = capi.default_params(), capi.default_step_params()
    assert (p.num_particles, p.grid_res, p.size) == (80000, 21, 1.0)          # Fluid.cpp:10-14
    assert p.particle_radius == np.float32(0.01) and p.time_scale == np.float32(0.012)
    assert (sp.viscosity_coefficient, sp.stiffness, sp.rest_density, sp.rest_pressure) == \
        (200.0, 100.0, 500.0, 0.0)                                            # Fluid.cpp:18-21
    assert list(sp.gravity) == [0.0, -900.0, 0.0]


@pytest.mark.parametrize("n,radius", [(80000, 0.01), (1_000_000, 0.01), (8_000_000, 0.0126)])
def test_derived_constants_equal_oracle(capi, oracle, n, radius):
    from watercube_b200 import scenes

    size, G = scenes.scaled_box(n, 0.01)
    d = capi.derive(capi.default_params(num_particles=n, size=size, grid_res=G,
                                        particle_radius=radius))
    o = oracle.derive(oracle.default_params(num_particles=n, size=size, grid_res=G,
                                            particle_radius=radius))
    for f in ("num_bins", "bin_size", "kernel_radius", "particle_mass", "poly6_const",
              "spiky_const", "visc_const"):
        assert getattr(d, f) == getattr(o, f), f
    # dist2_threshold: smallest fp32 x with sqrt(x) >= h
    h, T = np.float32(d.kernel_radius), np.float32(d.dist2_threshold)
    assert np.sqrt(T) >= h and np.sqrt(np.nextafter(T, np.float32(0))) < h


def test_invalid_arguments_are_rejected(capi):
    L = capi.lib()
    d = capi.Derived()
    assert L.wc_derive(C.byref(capi.default_params(grid_res=0)), C.byref(d)) == capi.WC_ERR_INVALID
    assert L.wc_derive(C.byref(capi.default_params(size=-1.0)), C.byref(d)) == capi.WC_ERR_INVALID
    assert b"grid_res" in L.wc_last_error() or b"size" in L.wc_last_error()
    assert L.wc_step(None, 0.016, None) == capi.WC_ERR_INVALID
    assert L.wc_sync(None) == capi.WC_ERR_INVALID


def test_no_cpu_fallback(capi):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.WcError) as e:
        capi.Fluid()
    assert e.value.code == capi.WC_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs may touch
    oracle/: not the product package, not the C-ABI header, not the dev tools."""
    for top in ("watercube_b200", "include", "tools"):
        for root, _, names in os.walk(os.path.join(ROOT, top)):
            for n in names:
                if n.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                    text = open(os.path.join(root, n), errors="replace").read()
                    assert "libwc_oracle" not in text, n
                    assert not re.search(r"#\s*include[^\n]*oracle", text), n
                    assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), n
