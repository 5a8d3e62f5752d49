"""In-tree build of the native library:  python -m watercube_b200.build

nvcc cross-compiles for sm_100a without a GPU.  The resulting
watercube_b200/csrc/libwc_sph.so is git-ignored but travels with the tree to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libwc_sph.so")
SOURCES = ["wc_capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--cudart", "static", "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _inputs():
    files = []
    for root, _, names in os.walk(CSRC):
        for n in names:
            if n.endswith((".cu", ".cuh", ".h")):
                files.append(os.path.join(root, n))
    files.append(os.path.join(os.path.dirname(CSRC), "..", "include", "wc_sph.h"))
    return files


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _inputs() if os.path.exists(f))


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines / out: build a tuning variant (e.g. defines=["WC_REPLAY_WORDS=10"]) elsewhere."""
    lib = out or LIB
    if not force and not out and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           *[f"-D{d}" for d in defines], "-o", lib, *[os.path.join(CSRC, s) for s in SOURCES]]
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that nvcc does not need; use the distro host compiler
    res = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(res.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
