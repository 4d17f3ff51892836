"""2-GPU check (torchrun) of the status word travelling inside the all-reduce: a structure in which one
unit needs the jitter rule (gpy_linalg.py:77-97) - the resident path reports it through the reduced status,
every rank repeats the evaluation synchronously, and the result equals the unsharded one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from gprf_b200 import GPRF, GPCov
from gprf_b200.dist import ShardedGPRF

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rng = np.random.RandomState(11)
base = rng.rand(12, 2)
Xa = np.repeat(base, 8, axis=0) + 1e-9 * rng.randn(96, 2)         # near-duplicates: needs jitter with nv < 0
Xb = rng.rand(200, 2)
X = np.vstack([Xa, Xb])
Y = rng.randn(len(X), 5)
blocks = [np.arange(0, 48), np.arange(48, 96), np.arange(96, 196), np.arange(196, 296)]
edges = [(1, 0), (3, 2)]
cov = GPCov([1.0], [0.3, 0.3], "euclidean", "se")
nv = -2e-4
gs = ShardedGPRF(X, Y, None, cov, nv, block_idxs=blocks, neighbors=edges)
ll, gX, gC = gs.llgrad(grad_X=True, grad_cov=True)
ev, fb, st = gs.resident_stats()
if rank == 0:
    g = GPRF(X, Y, None, cov, nv, block_idxs=blocks, neighbors=edges, device=lr)
    ll0, gX0, gC0 = g.llgrad(grad_X=True, grad_cov=True)
    print("jitter case world %d: ll %.6f vs %.6f, gradX rel %.2e, resident stats rank0 %s / unsharded %s"
          % (world, ll, ll0, np.abs(gX - gX0).max() / np.abs(gX0).max(), (ev, fb, st), g.resident_stats()), flush=True)
    assert abs(ll - ll0) <= 1e-9 * abs(ll0) and np.abs(gX - gX0).max() <= 1e-7 * np.abs(gX0).max()
    print("ok", flush=True)
    g.close()
# and an evaluation that fails on every jitter level raises on every rank, without a hang
gs.noise_var = -0.5
try:
    gs.llgrad()
    raised = False
except Exception as exc:       # noqa: BLE001
    raised = True
assert raised
if rank == 0:
    print("failure raised on every rank: ok", flush=True)
gs.noise_var = 0.05
gs.llgrad(grad_X=True)
gs.close()
dist.barrier()
dist.destroy_process_group()
