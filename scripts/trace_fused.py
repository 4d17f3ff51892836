"""Timeline of the fused per-unit kernel (debug aid, not a benchmark).

    python scripts/trace_fused.py [cfg2] [n_ctas] > gpurun_out/trace.txt

Enables gprf_debug_trace, runs a few device-resident evaluations and prints the (tag, time)
events thread 0 of each traced CTA recorded.  Tags: 1 start, 2 prep done, 31 diag product done,
32 diag tile in smem, 33 smem Cholesky+inverse done, 3 diag phase done, 41 panel product done,
42 panel triangular multiply done, 40 panel task done, 4 panel phase done, 50/5 trtri task/phase,
60/6 alpha task/phase, 70/7 kinv_grad task/phase (71 products done, 72/73 epilogue halves), 8 finalize done.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import bench

SLOTS = 512
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = bench.make_workload(name)
R = bench.Runner(torch, None, wl, 0, 1, 0)
g = R.g
g.set_fused_nt(8)            # the fused kernel is no longer the default schedule
sizes = bench.unit_sizes(wl["block_idxs"], wl["neighbors"])
order = np.argsort(-sizes, kind="stable")
nct = len(sizes)
for _ in range(3):
    R.device_step(reblock=R.reblock)
torch.cuda.synchronize()
g._check(g._lib.gprf_debug_trace(g._h, nct, None))
R.flush_l2()
R.device_step(reblock=R.reblock)
torch.cuda.synchronize()
buf = np.zeros((nct, SLOTS, 2), dtype=np.uint64)
g._check(g._lib.gprf_debug_trace(g._h, nct, buf.ctypes.data_as(C.c_void_p)))
tags = buf[:, :, 0].astype(np.int64)
t = buf[:, :, 1].astype(np.int64)
t0 = t[t > 0].min()
NAMES = {1: "start", 2: "prep", 31: "dg.gemm", 32: "dg.S", 33: "dg.chol", 3: "DIAG", 41: "pn.gemm", 42: "pn.mul",
         40: "pn.task", 4: "PANEL", 50: "tr.task", 5: "TRTRI", 60: "al.task", 6: "ALPHA", 71: "gr.gemm", 72: "gr.ep0", 73: "gr.ep1", 70: "gr.task", 7: "GRAD",
         8: "FIN"}
print("# units %d; kernel span %.1f us" % (nct, (t.max() - t0) / 1e3))
show = [0, 1, nct // 2, nct - 1] if len(sys.argv) < 3 else list(range(int(sys.argv[2])))
for c in show:
    k = int((t[c] > 0).sum())
    print("## cta %d  s=%d  start %.1f us  end %.1f us" % (c, int(sizes[order[c]]), (t[c, 0] - t0) / 1e3,
                                                          (t[c, k - 1] - t0) / 1e3))
    line = []
    for i in range(1, k):
        line.append("%s+%.1f" % (NAMES.get(int(tags[c, i]), str(tags[c, i])), (t[c, i] - t[c, i - 1]) / 1e3))
    print("   " + " ".join(line))
# aggregate: time attributed to each tag (interval ending at the tag) over units with 4 tiles
agg, cnt = {}, 0
for c in range(nct):
    if (int(sizes[order[c]]) + 63) // 64 != 4:
        continue
    k = int((t[c] > 0).sum())
    for i in range(1, k):
        n = NAMES.get(int(tags[c, i]), str(tags[c, i]))
        a = agg.setdefault(n, [0.0, 0])
        a[0] += (t[c, i] - t[c, i - 1]) / 1e3
        a[1] += 1
    cnt += 1
print("# per 4-tile unit (%d units): total us and mean us per event" % cnt)
for n, (v, m) in agg.items():
    print("#   %-8s %8.1f us/unit   %6.2f us x %.1f" % (n, v / cnt, v / m, m / cnt))
ends = np.array([t[c][t[c] > 0].max() - t0 for c in range(nct) if (t[c] > 0).any()]) / 1e3
starts = np.array([t[c][0] - t0 for c in range(nct) if (t[c] > 0).any()]) / 1e3
print("# CTA start times: %d at <5us, max %.1f; end times pct 50/90/100: %.1f %.1f %.1f" % (
    (starts < 5).sum(), starts.max(), np.percentile(ends, 50), np.percentile(ends, 90), ends.max()))
