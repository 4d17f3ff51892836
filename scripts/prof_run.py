"""Minimal driver for ncu: a few device-resident evaluations of one bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.make_workload(name)
R = bench.Runner(torch, None, wl, 0, 1, 0)
for _ in range(n):
    R.device_step(reblock=R.reblock)
torch.cuda.synchronize()
print("done", name, R.g.last_timing())
