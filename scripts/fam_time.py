"""Per-family device time of one workload (shards emulated on one GPU).  python scripts/fam_time.py cfg5 [world...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gprf_b200 import GPRF

wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg5")
worlds = [int(a) for a in sys.argv[2:]] or [1]
dev = torch.device("cuda", 0)
n, dx = wl["X"].shape
Xd = torch.tensor(wl["X"], dtype=torch.float64, device=dev)
out = torch.zeros(1 + 5 + n * dx, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for world in worlds:
    g = GPRF(wl["X"], wl["Y"], wl["block_fn"], wl["cov"], wl["noise_var"], block_idxs=wl["block_idxs"],
             neighbors=wl["neighbors"], device=0, neighbor_threshold=wl.get("threshold", 1e-3),
             unit_shard=(0, world) if world > 1 else None)
    if wl["neighbors"] is None:
        wl["neighbors"] = list(g.neighbors)
    if os.environ.get("REUSE"):
        g.set_factor_reuse(int(os.environ["REUSE"]))
    print("blocks %d edges %d reuse %s" % (g.n_blocks, len(g.neighbors), g.factor_reuse_stats()), flush=True)
    ts = []
    for it in range(5):
        if it == 3:
            g.set_profiling(True)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.llgrad_device(Xd.data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream,
                        grad_X=True, grad_cov=wl["grad_cov"], reblock=g._device_part is not None)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    fam = g.family_timing()
    print("world %d: %.3f ms/eval (min of unprofiled %s)  %s" % (
        world, min(ts[1:3]), ["%.2f" % t for t in ts],
        " ".join("%s=%.2f" % (k, v[0]) for k, v in fam.items() if v[0] > 0.005)), flush=True)
    g.close()
