"""Offline model of the resident launch's static unit lists (resident.cuh: res_plan_body) on MEASURED unit times.

    python scripts/plan_sim.py [profiles/r02k_unit_times_cfg2.txt]

The input is the per-unit table `scripts/trace_resident.py cfg2 <file>` writes (uid, points of block i, points
of block j, us, CTA, start us).  A cost model is fitted to it (pair = c1 ab + c2 bb + c0, block = d1 bb + d0, in
8-point blocks) and three assignments of the units to 148 CTAs are replayed with the parents' release times:
  cur      the shipped plan: blocks in id order, pairs by cost key descending, the CTAs that take one pair more
           get the smallest pairs, snake order inside each class
  sortblk  the same with the block units dealt in size order (smallest next to the largest pairs) - shipped
  match    sortblk + the partner of each top pair chosen as the free pair closest to the level load - not built
Every line: longest CTA, mean of the block-carrying CTAs, longest / mean of the others, with 10 us of noise on the
pair times (the fit's residual) and for resampled block sizes.  This is the evidence behind DESIGN.md section 9.1.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = 148
WAIT = 40.0          # a CTA without a block unit: its first pair waits for the parent's W (us)
OWN = 8.6            # ... of which the pair's own gather + K_ji fill this much (deferred W_i request)


def load(path):
    r = np.loadtxt(path)
    pairs = r[r[:, 1] > 0]
    blocks = r[r[:, 1] == 0]
    A = np.c_[np.ceil(pairs[:, 1] / 8), np.ceil(pairs[:, 2] / 8), np.ones(len(pairs))]
    cp = np.linalg.lstsq(A, pairs[:, 3], rcond=None)[0]
    rms = float(np.sqrt(np.mean((A @ cp - pairs[:, 3]) ** 2)))
    cb = np.polyfit(np.ceil(blocks[:, 2] / 8), blocks[:, 3], 1)
    B = len(blocks)
    sizes = np.zeros(B, dtype=int)
    sizes[blocks[:, 0].astype(int)] = blocks[:, 2].astype(int)
    # edges (i, j) from the sizes are ambiguous; keep (a, b) per pair instead and the parent's id unknown: the
    # release time of a parent is modelled by its size alone
    return sizes, pairs[:, 1:3].astype(int), cp, cb, rms


def replay(lists, bcost, pcost, prel):
    ends = []
    for units in lists:
        t = 0.0
        for kind, u in units:
            if kind == "b":
                t += bcost[u]
            else:
                t = max(t + OWN, prel[u]) + pcost[u] - OWN
        ends.append(t)
    return np.array(ends)


def plan(key, bsz, mode):
    E, B = len(key), len(bsz)
    order = sorted(range(E), key=lambda e: (-key[e], e))
    border = sorted(range(B), key=lambda b: ((bsz[b] + 7) // 8, b)) if mode != "cur" else list(range(B))
    lists = [[] for _ in range(G)]
    for k, b in enumerate(border):
        lists[k % G].append(("b", b))
    q, rem = E // G, E % G
    mB = G - rem
    if mode == "match" and q == 2:
        cost = 3.8 * np.array([key[e] for e in order]) - 225.0
        bm = 46.0 + 2.5 * np.array([(bsz[b] + 7) // 8 for b in border])
        T = (cost.sum() + bm.sum() + WAIT * (G - min(B, G))) / G
        free = np.ones(E, bool)
        free[:mB] = False
        for w in range(mB):
            own = bm[w] if w < B else WAIT
            cand = np.where(free)[0]
            j = cand[np.argmin(np.abs(cost[cand] - (T - own - cost[w])))]
            free[j] = False
            lists[w] += [("p", order[w]), ("p", order[j])]
        for idx, r in enumerate(np.where(free)[0]):
            rnd, pos = divmod(idx, rem)
            lists[mB + ((rem - 1 - pos) if rnd & 1 else pos)].append(("p", order[r]))
        return lists
    for rank, e in enumerate(order):
        if rank < mB * q:
            rnd, pos = divmod(rank, mB)
            w = (mB - 1 - pos) if rnd & 1 else pos
        else:
            rnd, pos = divmod(rank - mB * q, rem)
            w = mB + ((rem - 1 - pos) if rnd & 1 else pos)
        lists[w].append(("p", e))
    return lists


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02k_unit_times_cfg2.txt")
    sizes, ab, cp, cb, rms = load(path)
    print("pair us = %.2f ab + %.2f bb %+.1f (rms %.1f);  block us = %.2f bb + %.1f" % (cp[0], cp[1], cp[2], rms, cb[0], cb[1]))
    rng = np.random.RandomState(5)
    E = len(ab)
    for trial in range(6):
        if trial == 0:
            a8, b8, bsz = np.ceil(ab[:, 0] / 8), np.ceil(ab[:, 1] / 8), sizes
        else:                                   # the same structure with resampled block sizes
            scale = rng.multinomial(sizes.sum(), sizes / sizes.sum()) / np.maximum(sizes, 1)
            f = scale[rng.randint(0, len(sizes), size=(E, 2))]
            a8, b8 = np.ceil(ab[:, 0] * f[:, 0] / 8), np.ceil(ab[:, 1] * f[:, 1] / 8)
            bsz = (sizes * scale).astype(int)
        key = (3 * a8 + 5 * b8).astype(int)
        pcost = cp[0] * a8 + cp[1] * b8 + cp[2] + rng.randn(E) * rms
        bcost = cb[0] * np.ceil(bsz / 8) + cb[1]
        prel = 0.58 * (cb[0] * a8 + cb[1])      # the parent's W is exported at ~58 % of its block unit
        out = []
        for mode in ("cur", "sortblk", "match"):
            e = replay(plan(key, bsz, mode), bcost, pcost, prel)
            nb = min(len(bsz), G)
            out.append("%s max %.0f (blk mean %.0f | other max %.0f mean %.0f)" % (mode, e.max(), e[:nb].mean(), e[nb:].max(), e[nb:].mean()))
        print("%s  %s" % ("measured sizes " if trial == 0 else "resampled %d    " % trial, "  ".join(out)))


if __name__ == "__main__":
    main()
