"""A/B of the edge factor reuse (gprf_set_factor_reuse) on the n=200k workload: per-family device
time with the reuse off and on, the results compared bit for bit, and the per-rank time of the
group-LPT shards emulated on one GPU.   python scripts/reuse_ab.py [worlds...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from gprf_b200 import GPRF

wl = bench.make_workload(os.environ.get("WL", "cfg5"))
dev = torch.device("cuda", 0)
n, dx = wl["X"].shape
Xd = torch.tensor(wl["X"], dtype=torch.float64, device=dev)
out = torch.zeros(1 + 5 + n * dx, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
worlds = [int(a) for a in sys.argv[1:]] or [1]


def run(g, reps=4):
    ts = []
    for it in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.llgrad_device(Xd.data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream,
                        grad_X=True, grad_cov=False, reblock=g._device_part is not None)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), out.clone()


for world in worlds:
    for rank in sorted(set([0, world - 1])):
        g = GPRF(wl["X"], wl["Y"], wl["block_fn"], wl["cov"], wl["noise_var"], block_idxs=wl["block_idxs"],
                 neighbors=wl["neighbors"], device=0, unit_shard=(rank, world) if world > 1 else None)
        res = {}
        for on in (False, True):
            g.set_factor_reuse(on)
            ms, o = run(g)
            res[on] = o
            g.set_profiling(True)
            run(g, 2)
            fam = g.family_timing()
            g.set_profiling(False)
            print("world %d rank %d reuse %-5s: %.2f ms/eval  stats %s  %s" % (
                world, rank, on, ms, g.factor_reuse_stats(),
                " ".join("%s=%.2f" % (k, v[0]) for k, v in fam.items() if v[0] > 0.005)), flush=True)
        print("   bit-identical:", bool(torch.equal(res[False], res[True])), flush=True)
        g.close()
