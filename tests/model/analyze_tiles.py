"""Offline model of the warp-cooperative gather: candidates scanned / surviving the box cull
per 32-target warp, for a cell-sorted scene.  Used to choose grouping and cull strategy
without spending GPU time."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as ob
from watercube_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
group = int(sys.argv[3]) if len(sys.argv) > 3 else 32
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
pos = ob.as_f32(s["sorted"])[:, :3]
G, h, T = p.grid_res, d.kernel_radius, d.kernel_radius ** 2
offsets = np.concatenate([s["offsets"], [n]]).astype(np.int64)
cid = s["cell_ids"][s["perm"]].astype(np.int64)
cx, cy, cz = cid % G, (cid // G) % G, cid // (G * G)
row = cz * G + cy
rng = np.random.default_rng(0)
warps = rng.choice(n // group, size=min(400, n // group), replace=False)
tot_scan = tot_surv = tot_pass = tot_true = tot_tgt = 0
tot_surv_tight = 0
for w in warps:
    lo, hi = w * group, min(n, (w + 1) * group)
    for r in np.unique(row[lo:hi]):
        sel = np.arange(lo, hi)[row[lo:hi] == r]
        tp = pos[sel]
        x0, x1 = max(cx[sel].min() - 1, 0), min(cx[sel].max() + 1, G - 1)
        ry, rz = r % G, r // G
        b0, b1 = tp.min(0), tp.max(0)
        cand = []
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                z, y = rz + dz, ry + dy
                if 0 <= z < G and 0 <= y < G:
                    base = (z * G + y) * G
                    cand.append(np.arange(offsets[base + x0], offsets[base + x1 + 1]))
        cand = np.concatenate(cand)
        q = pos[cand]
        e = np.maximum(np.maximum(b0 - q, q - b1), 0)
        surv = (e * e).sum(1) < T
        d2 = ((tp[:, None, :] - q[None, surv, :]) ** 2).sum(-1)
        tot_scan += len(cand); tot_surv += surv.sum(); tot_pass += 1
        tot_true += (d2 < T).sum() - len(sel); tot_tgt += len(sel)
        tot_surv_tight += (d2 < T).any(0).sum()
nw = len(warps)
print(f"n={n} steps={steps} group={group}: passes/warp={tot_pass/nw:.2f} scanned/warp={tot_scan/nw:.0f} "
      f"survivors/warp={tot_surv/nw:.0f} union-of-true/warp={tot_surv_tight/nw:.0f} "
      f"true-neigh/target={tot_true/tot_tgt:.1f} particles/cell={n/(np.diff(offsets)>0).sum():.1f}")
