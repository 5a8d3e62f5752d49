"""Offline model of the density pass's test count per target (no GPU): what the kept design
tests (every candidate within h of the bounding box of a row-aligned 32-target group) against two
designs that would test fewer -- per-lane cell-column windows inside the same groups, and
row-aligned 16-target groups -- and against the true neighbour count.  DESIGN.md "where the next
factor would come from" quotes these numbers.

    python tests/model/model_tests.py [particles] [steps before measuring]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as ob            # noqa: E402  (tools may use the checker)
from watercube_b200 import scenes           # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = scenes.dam_break(n)
p = ob.default_params(num_particles=sc.n, size=sc.size, grid_res=sc.grid_res)
d = ob.derive(p)
P = sc.particles
if steps:
    st = ob.Stepper(P, p, nthreads=ob.max_threads())
    for _ in range(steps):
        st.step(1 / 60)
    P = ob.as_f32(st.buf1).copy()
s = ob.sort(P, d.bin_size, p.grid_res)
pos = ob.as_f32(s["sorted"])[:, :3].astype(np.float64)
G, T = p.grid_res, float(d.kernel_radius) ** 2
offsets = np.concatenate([s["offsets"], [n]]).astype(np.int64)
cid = s["cell_ids"][s["perm"]].astype(np.int64)
cx, row = cid % G, cid // G
row_start = offsets[::G][:G * G]
row_end = offsets[G::G][:G * G]


def groups_of(size):
    out = []
    for r in np.flatnonzero(row_end > row_start):
        for b in range(row_start[r], row_end[r], size):
            out.append((r, b, min(b + size, row_end[r])))
    return out


def slices(r, x0, x1):
    ry, rz = r % G, r // G
    idx, col = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            z, y = rz + dz, ry + dy
            if 0 <= z < G and 0 <= y < G:
                base = (z * G + y) * G
                a = np.arange(offsets[base + x0], offsets[base + x1 + 1])
                idx.append(a)
                col.append(cx[a])
    return np.concatenate(idx), np.concatenate(col)


def measure(size, sample=600, seed=0):
    gs = groups_of(size)
    rng = np.random.default_rng(seed)
    pick = rng.choice(len(gs), size=min(sample, len(gs)), replace=False)
    streamed = kept = window = true = targets = 0
    for k in pick:
        r, lo, hi = gs[k]
        tp = pos[lo:hi]
        tc = cx[lo:hi]
        x0, x1 = max(tc.min() - 1, 0), min(tc.max() + 1, G - 1)
        cand, ccol = slices(r, x0, x1)
        q = pos[cand]
        e = np.maximum(np.maximum(tp.min(0) - q, q - tp.max(0)), 0)
        keep = (e * e).sum(1) < T                   # the kept design's cull
        d2 = ((tp[:, None, :] - q[None, keep, :]) ** 2).sum(-1)
        streamed += len(cand)
        kept += keep.sum() * len(tp)                # every lane tests every kept candidate
        window += (np.abs(ccol[keep][None, :] - tc[:, None]) <= 1).sum()   # own column +- 1 only
        true += (d2 < T).sum() - len(tp)
        targets += len(tp)
    return dict(groups=len(gs), lanes_used=targets / (len(pick) * size), streamed_per_group=streamed / len(pick),
                tests_per_target=kept / targets, window_tests_per_target=window / targets,
                neighbours=true / targets)


for size in (32, 16):
    m = measure(size)
    print(f"n={n} steps={steps} group<={size}: {m['groups']} groups, lane use {m['lanes_used']:.2f}, "
          f"streamed/group {m['streamed_per_group']:.0f}, tests/target {m['tests_per_target']:.0f}, "
          f"with column windows {m['window_tests_per_target']:.0f}, true neighbours {m['neighbours']:.1f}")
