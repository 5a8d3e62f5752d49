"""Build the C++ host facade (libwc_core.so) and the headless driver:
python -m watercube_b200.host.build   (needs csrc/libwc_sph.so: python -m watercube_b200.build)"""
from __future__ import annotations

import os
import subprocess
import sys

HOST = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HOST, "libwc_core.so")
EXE = os.path.join(HOST, "wc_headless")


def build(force: bool = False) -> str:
    from .. import build as native

    native.build()
    cmd = ["make", "-C", HOST] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("building the host facade failed")
    return EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
