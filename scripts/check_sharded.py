"""N-GPU check (torchrun): ShardedGPRF (group-LPT shards, factor reuse, one NCCL all-reduce) against
the unsharded evaluation of the same problem on rank 0's GPU.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from gprf_b200 import GPRF
from gprf_b200.dist import ShardedGPRF

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for name in sys.argv[1:] or ["cfg1", "cfg5"]:
    wl = bench.make_workload(name)
    kw = dict(block_idxs=wl["block_idxs"], neighbors=wl["neighbors"])
    gs = ShardedGPRF(wl["X"], wl["Y"], wl["block_fn"], wl["cov"], wl["noise_var"], **kw)
    X2 = wl["X"] + 1e-4 * np.random.RandomState(1).randn(*wl["X"].shape)
    gs.update_X(X2)
    ll, gX, gC = gs.llgrad(grad_X=True, grad_cov=True)
    stats = gs.factor_reuse_stats()
    if rank == 0:
        g = GPRF(wl["X"], wl["Y"], wl["block_fn"], wl["cov"], wl["noise_var"], device=lr, **kw)
        g.update_X(X2)
        ll0, gX0, gC0 = g.llgrad(grad_X=True, grad_cov=True)
        print("%s world %d: ll %.6f vs %.6f rel %.2e | gradX rel %.2e | gradCov rel %.2e | reuse on rank0 %s / unsharded %s"
              % (name, world, ll, ll0, abs(ll - ll0) / abs(ll0), np.abs(gX - gX0).max() / np.abs(gX0).max(),
                 np.abs(gC - gC0).max() / np.abs(gC0).max(), stats, g.factor_reuse_stats()), flush=True)
        assert abs(ll - ll0) <= 1e-12 * abs(ll0) and np.abs(gX - gX0).max() <= 1e-11 * np.abs(gX0).max()
        g.close()
    gs.close()
    dist.barrier()
dist.destroy_process_group()
