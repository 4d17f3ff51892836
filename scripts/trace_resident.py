"""Timeline of the resident unit kernel (debug aid, not a benchmark).

    python scripts/trace_resident.py [cfg2] > gpurun_out/trace_resident.txt

Thread 0 of every CTA records (unit, tag, %globaltimer) at the phase boundaries of resident.cuh:
1 start, 2 P1 L_ji done, 3 P2 S done, 4 Cholesky done, 5 inverse done, 6 Y part (Z, alpha_j) done,
7 T done, 8 V done, 9 alpha_i done, 10 G contraction done, 11 finalize done.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402

SLOTS = 512
NAMES = {2: "P1 L_ji", 3: "P2 S", 4: "P3a chol", 5: "P3b inv", 6: "P4 Y-part", 7: "P6a T", 8: "P6b V",
         9: "P7 alpha_i", 10: "P8 G/grad", 11: "finalize", 19: "P8 piece barrier", 20: "P8 table+issue", 21: "P8 tma wait",
         22: "P8 warp0 tasks", 30: "chol: barrier wait", 31: "chol: panel + diag update", 32: "chol: diag block", 33: "gather", 34: "P1 cov K_ji", 35: "P1 wait W_i", 36: "P4 R=Y-LZ", 37: "P4 Z=W R + stores", 40: "V: warp 0 products", 41: "V: barrier wait", 42: "T: wait for W_i (TMA)", 43: "T: warp 0 products",
         44: "T: barrier wait", 45: "S: entry", 46: "S: warp 0 tasks"}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = bench.make_workload(name)
R = bench.Runner(torch, None, wl, 0, 1, 0)
g = R.g
nct = 148
for _ in range(3):
    R.device_step(reblock=R.reblock)
torch.cuda.synchronize()
print("resident stats", g.resident_stats())
g._check(g._lib.gprf_debug_trace(g._h, nct, None))
R.flush_l2()
R.device_step(reblock=R.reblock)
torch.cuda.synchronize()
buf = np.zeros((nct, SLOTS, 2), dtype=np.uint64)
g._check(g._lib.gprf_debug_trace(g._h, nct, buf.ctypes.data_as(C.c_void_p)))
tag = (buf[:, :, 0] & np.uint64(0xffff)).astype(np.int64)
uid = (buf[:, :, 0] >> np.uint64(16)).astype(np.int64)
t = buf[:, :, 1].astype(np.int64)
B = len(wl["block_idxs"])
t0 = t[t > 0].min()
print("# span of all marks %.1f us" % ((t.max() - t0) / 1e3))
agg = {"block": {}, "pair": {}}
cnt = {"block": 0, "pair": 0}
tot = {"block": 0.0, "pair": 0.0}
for c in range(nct):
    k = int((t[c] > 0).sum())
    for i in range(1, k):
        if tag[c, i] == 1 or uid[c, i] != uid[c, i - 1]:
            continue
        kind = "block" if uid[c, i] < B else "pair"
        d = (t[c, i] - t[c, i - 1]) / 1e3
        agg[kind].setdefault(int(tag[c, i]), []).append(d)
        tot[kind] += d
    for i in range(k):
        if tag[c, i] == 1:
            cnt["block" if uid[c, i] < B else "pair"] += 1
for kind in ("block", "pair"):
    if not cnt[kind]:
        continue
    print("## %s units traced: %d, mean total %.1f us" % (kind, cnt[kind], tot[kind] / cnt[kind]))
    for tg in sorted(agg[kind]):
        v = np.array(agg[kind][tg])
        print("#   %-12s mean %7.2f us   min %7.2f   max %7.2f   (%d)" % (NAMES.get(tg, str(tg)), v.mean(), v.min(),
                                                                         v.max(), len(v)))
# per-CTA busy span of the last launch (pairs)
ends = []
for c in range(nct):
    k = int((t[c] > 0).sum())
    if k:
        ends.append((t[c, k - 1] - t0) / 1e3)
        first_pair = [i for i in range(k) if uid[c, i] >= B]
        if c < 3 and first_pair:
            print("# cta %d: units %s" % (c, sorted(set(uid[c, :k].tolist()))))
ends = np.array(ends)
print("# CTA end times pct 0/50/90/100: %.1f %.1f %.1f %.1f us" % (ends.min(), np.percentile(ends, 50),
                                                                   np.percentile(ends, 90), ends.max()))
# per-unit totals with the unit's sizes (a = parent block, b = own block), for the plan's cost model
if len(sys.argv) > 2:
    blocks = g.block_idxs
    sizes = [len(b) for b in blocks]
    edges = wl["neighbors"]
    rows = []
    for c in range(nct):
        k = int((t[c] > 0).sum())
        start = {}
        for i in range(k):
            u = int(uid[c, i])
            if tag[c, i] == 1:
                start[u] = t[c, i]
            elif tag[c, i] == 11 and u in start:
                if u < B:
                    rows.append((u, 0, sizes[u], (t[c, i] - start[u]) / 1e3, c, (start[u] - t0) / 1e3))
                else:
                    i_, j_ = edges[u - B]
                    rows.append((u, sizes[i_], sizes[j_], (t[c, i] - start[u]) / 1e3, c, (start[u] - t0) / 1e3))
    np.savetxt(sys.argv[2], np.array(rows), fmt="%.3f", header="uid a b us cta start_us")
