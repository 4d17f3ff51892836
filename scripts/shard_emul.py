"""Per-rank device time of the n=200k workload when its units are sharded over W ranks,
emulated on ONE GPU (rank r's shard only, no all-reduce): shows how the strong-scaling
shards behave without needing W GPUs.   python scripts/shard_emul.py [fused_nt ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from gprf_b200 import GPRF

wl = bench.make_workload("cfg5")
dev = torch.device("cuda", 0)
n, dx = wl["X"].shape
Xd = torch.tensor(wl["X"], dtype=torch.float64, device=dev)
out = torch.zeros(1 + 5 + n * dx, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fnts = [int(a) for a in sys.argv[1:]] or [8, 20]
for world in (1, 2, 4, 8):
    for rank in sorted(set([0, world - 1])):
        g = GPRF(wl["X"], wl["Y"], wl["block_fn"], wl["cov"], wl["noise_var"], block_idxs=wl["block_idxs"],
                 neighbors=wl["neighbors"], device=0, unit_shard=(rank, world) if world > 1 else None)
        for fnt in fnts:
            g.set_fused_nt(fnt)
            ts = []
            for it in range(4):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.llgrad_device(Xd.data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream,
                                grad_X=True, grad_cov=False, reblock=g._device_part is not None)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            print("world %d rank %d fused_nt %2d: %.2f ms/eval (min of %s)" % (
                world, rank, fnt, min(ts[1:]), ["%.1f" % t for t in ts]), flush=True)
        g.close()
