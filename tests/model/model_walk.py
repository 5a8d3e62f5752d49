"""Offline model of the update pass's replay walk (wc_sph_tile.cuh: UpdateAcc::walk): pair-loop
iterations per 32-target group for different ways of batching the neighbour-list words.
Every lane drains its own accepted bits two per iteration and lanes only wait for each other
at a batch's end, so iterations/batch = max over lanes of ceil(bits / 2).

    python tests/model/model_walk.py [n] [steps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as ob  # noqa: E402
from watercube_b200 import scenes  # noqa: E402

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
pos = ob.as_f32(s["sorted"])[:, :3]
G, T = p.grid_res, np.float32(d.kernel_radius) ** 2
offsets = np.concatenate([s["offsets"], [n]]).astype(np.int64)
cid = s["cell_ids"][s["perm"]].astype(np.int64)
cx = cid % G

# groups: every (y,z) row cut into runs of <= 32 consecutive sorted particles
groups = []
for r in range(G * G):
    b, e = offsets[r * G], offsets[(r + 1) * G]
    for k in range(b, e, 32):
        groups.append((k, min(32, e - k), r))
rng = np.random.default_rng(0)
pick = rng.choice(len(groups), size=min(600, len(groups)), replace=False)


def iters(bits_per_word, batches):
    """bits_per_word: [words, lanes] accepted counts; batches: list of word-index arrays."""
    it = 0
    for ws in batches:
        c = bits_per_word[ws].sum(0)
        it += int(np.ceil(c.max() / 2))
    return it


res = {}
words_tot = surv_tot = pairs_tot = lanes_tot = 0
for gi in pick:
    first, cnt, r = groups[gi]
    sel = np.arange(first, first + cnt)
    tp = pos[sel]
    x0, x1 = max(cx[sel].min() - 1, 0), min(cx[sel].max() + 1, G - 1)
    ry, rz = r % G, r // G
    b0, b1 = tp.min(0), tp.max(0)
    cand = []
    for sl in range(9):
        z, y = rz + sl // 3 - 1, ry + sl % 3 - 1
        if 0 <= z < G and 0 <= y < G:
            base = (z * G + y) * G
            cand.append(np.arange(offsets[base + x0], offsets[base + x1 + 1]))
    cand = np.concatenate(cand)
    q = pos[cand]
    e = np.maximum(np.maximum(b0 - q, q - b1), 0)
    surv = cand[(e * e).sum(1) < T * 1.0001]
    q = pos[surv]
    acc = ((tp[:, None, :] - q[None, :, :]) ** 2).sum(-1) < T       # [lanes, cands]
    acc &= surv[None, :] != sel[:, None]                             # self pair dropped
    nw = (len(surv) + 31) // 32
    pad = np.zeros((cnt, nw * 32), bool)
    pad[:, :len(surv)] = acc
    bpw = pad.reshape(cnt, nw, 32).sum(-1).T                         # [words, lanes]
    words_tot += nw
    surv_tot += len(surv)
    pairs_tot += acc.sum()
    lanes_tot += cnt
    for K in (4, 5, 7, 10, 14, 99):
        nb = (nw + K - 1) // K
        cons = [np.arange(b * K, min(nw, (b + 1) * K)) for b in range(nb)]
        strd = [np.arange(b, nw, nb) for b in range(nb)]
        res.setdefault(("consecutive", K), []).append(iters(bpw, cons))
        res.setdefault(("strided", K), []).append(iters(bpw, strd))
    # lower bound: perfect balance within the group
    res.setdefault(("ideal", 0), []).append(int(np.ceil(acc.sum(1).max() / 2)))
    res.setdefault(("mean-lane", 0), []).append(acc.sum(1).mean() / 2)

ng = len(pick)
print(f"n={n} steps={steps}: groups={len(groups)} words/group={words_tot/ng:.1f} "
      f"survivors/group={surv_tot/ng:.0f} pairs/lane={pairs_tot/lanes_tot:.1f}")
for k in sorted(res):
    print(f"  {k[0]:12s} K={k[1]:3d}: {np.mean(res[k]):6.1f} iterations/group")


# ---- sliding ring: R resident words, refilled g words at a time once every lane has left them
def ring_iters(bits_lane_word, R, g):
    """bits_lane_word: [lanes, words] accepted counts.  Returns pair-loop iterations when a lane
    may run ahead of the slowest lane by the resident window."""
    lanes, nw = bits_lane_word.shape
    rem = bits_lane_word.astype(np.int64).copy()
    w = np.zeros(lanes, np.int64)             # word each lane is draining
    base, it = 0, 0                           # resident words: [base, base + R)
    for ln in range(lanes):
        while w[ln] < nw and rem[ln, w[ln]] == 0:
            w[ln] += 1
    while (w < nw).any():
        it += 1
        for ln in range(lanes):
            for _ in range(2):
                while w[ln] < min(nw, base + R) and rem[ln, w[ln]] == 0:
                    w[ln] += 1
                if w[ln] < min(nw, base + R):
                    rem[ln, w[ln]] -= 1
            while w[ln] < min(nw, base + R) and rem[ln, w[ln]] == 0:
                w[ln] += 1
        wmin = w.min()
        while base + g <= wmin:
            base += g
        if it > 10000:
            raise RuntimeError("stuck")
    return it


ring = {}
for gi in pick[:200]:
    first, cnt, r = groups[gi]
    sel = np.arange(first, first + cnt)
    tp = pos[sel]
    x0, x1 = max(cx[sel].min() - 1, 0), min(cx[sel].max() + 1, G - 1)
    ry, rz = r % G, r // G
    b0, b1 = tp.min(0), tp.max(0)
    cand = []
    for sl in range(9):
        z, y = rz + sl // 3 - 1, ry + sl % 3 - 1
        if 0 <= z < G and 0 <= y < G:
            base_ = (z * G + y) * G
            cand.append(np.arange(offsets[base_ + x0], offsets[base_ + x1 + 1]))
    cand = np.concatenate(cand)
    q = pos[cand]
    e = np.maximum(np.maximum(b0 - q, q - b1), 0)
    surv = cand[(e * e).sum(1) < T * 1.0001]
    q = pos[surv]
    acc = ((tp[:, None, :] - q[None, :, :]) ** 2).sum(-1) < T
    acc &= surv[None, :] != sel[:, None]
    nw = (len(surv) + 31) // 32
    pad = np.zeros((cnt, nw * 32), bool)
    pad[:, :len(surv)] = acc
    blw = pad.reshape(cnt, nw, 32).sum(-1)
    for R, g in ((4, 1), (5, 1), (6, 1), (6, 2), (6, 3), (8, 1), (8, 2), (8, 4), (10, 5), (5, 5), (14, 14)):
        ring.setdefault((R, g), []).append(ring_iters(blw, R, g))
for k in sorted(ring):
    print(f"  ring R={k[0]:2d} refill-granularity={k[1]:2d}: {np.mean(ring[k]):6.1f} iterations/group")
