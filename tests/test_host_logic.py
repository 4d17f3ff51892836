"""CPU-side checks: partitioners vs the oracle, the C-ABI surface, sharding algebra."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blocking_matches_oracle_bit_exact():
    from gprf_b200 import blocking as prod
    from oracle import blocking as orc
    rng = np.random.RandomState(0)
    X = rng.rand(5000, 2)
    for nb in (4, 20, 100):
        cp, co = prod.grid_centers(nb), orc.grid_centers(nb)
        assert all(np.array_equal(a, b) for a, b in zip(cp, co))
        bp, bo = prod.Blocker(cp), orc.Blocker(co)
        assert bp.neighbors() == bo.neighbors()
        assert bp.neighbors(False) == bo.neighbors(False)
        for a, b in zip(bp.block_clusters(X), bo.block_clusters(X)):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    # duplicate / equidistant points exercise the first-index tie break
    Xt = np.array([[0.5, 0.5], [0.25, 0.5], [0.5, 0.25], [0.0, 0.0], [1.0, 1.0]])
    for a, b in zip(prod.Blocker(prod.grid_centers(4)).block_clusters(Xt),
                    orc.Blocker(orc.grid_centers(4)).block_clusters(Xt)):
        assert np.array_equal(a, b)


def test_pdtree_and_rpc_match_oracle():
    from gprf_b200 import blocking as prod
    from oracle import blocking as orc
    rng = np.random.RandomState(4)
    X = np.column_stack([rng.uniform(-40, 340, 3000), rng.uniform(-70, 70, 3000), rng.rand(3000) * 300])
    ip, rp = prod.pdtree_cluster(X, blocksize=210)
    io, ro = orc.pdtree_cluster(X, blocksize=210)
    assert len(ip) == len(io) and all(np.array_equal(a, b) for a, b in zip(ip, io))
    Xm = X + rng.randn(*X.shape) * [0.2, 0.2, 20]
    Xc = Xm.copy()
    assert all(np.array_equal(a, b) for a, b in zip(rp(Xm), ro(Xm)))
    assert np.array_equal(Xm, Xc)
    np.random.seed(3)
    cp, sp = prod.cluster_rpc(X[:, :2], np.arange(3000), 300)
    np.random.seed(3)
    co, so = orc.cluster_rpc(X[:, :2], np.arange(3000), 300)
    assert len(cp) == len(co) and all(np.array_equal(a, b) for a, b in zip(cp, co))
    c2 = prod.cluster_rpc(Xm[:, :2], np.arange(3000), 300, fixed_split=sp)[0]
    o2 = orc.cluster_rpc(Xm[:, :2], np.arange(3000), 300, fixed_split=so)[0]
    assert all(np.array_equal(a, b) for a, b in zip(c2, o2))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads (no GPU needed) and exports what include/gprf_b200.h declares."""
    from gprf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    header = open(os.path.join(ROOT, "include", "gprf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gprf_[a-z_0-9]+)\s*\(", header)))
    assert declared == _lib.EXPORTS
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.gprf_abi_version() == 1
    assert lib.gprf_strerror(1) == b"not positive definite, even with jitter."
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert (" T " + name) in out


def test_shard_units_partitions_and_balances():
    from gprf_b200.dist import shard_units, unit_costs
    rng = np.random.RandomState(1)
    sizes = rng.randint(60, 140, size=100)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    from oracle.blocking import Blocker, grid_centers
    edges = np.array(Blocker(np.asarray(grid_centers(100))).neighbors())
    cost = unit_costs(ptr, edges)
    for world in (1, 2, 4, 8):
        masks = [shard_units(ptr, edges, r, world) for r in range(world)]
        assert np.array_equal(np.sum(masks, axis=0), np.ones(442))
        loads = [cost[m.astype(bool)].sum() for m in masks]
        assert max(loads) / (sum(loads) / world) < 1.05
        # every pair sits on the rank of its parent block i (factor reuse needs it there)
        for m in masks:
            assert np.array_equal(m[100:], m[edges[:, 0]])


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gprf_b200.dist import shard_units, pack, unpack
    from gprf_b200.gprf import _blocks_to_csr
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_oracle_props import small_gprf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gp = small_gprf("euclid_se", n=90, nb=5)
    n, dx = gp.X.shape
    ptr, _ = _blocks_to_csr(gp.block_idxs)
    mask = shard_units(ptr, np.asarray(gp.neighbors), rank, world).astype(bool)
    B = gp.n_blocks
    # partial sums over this rank's units, with the weights of gprf.py:253-291
    ll, gX, gC = 0.0, np.zeros((n, dx)), np.zeros(4)
    for b in range(B):
        if mask[b]:
            l, x, c = gp.llgrad_unary(b, grad_X=True, grad_cov=True)
            w = 1 - gp.neighbor_count[b]
            ll += w * l
            gX[gp.block_idxs[b]] += w * x
            gC += w * c
    for e, (i, j) in enumerate(gp.neighbors):
        if mask[B + e]:
            l, x, c = gp.llgrad_joint(i, j, grad_X=True, grad_cov=True)
            ni = len(gp.block_idxs[i])
            ll += l
            gX[gp.block_idxs[i]] += x[:ni]
            gX[gp.block_idxs[j]] += x[ni:]
            gC += c
    t = torch.from_numpy(pack(ll, gX, gC, n, dx))
    dist.all_reduce(t)
    got = unpack(t.numpy(), n, dx, 4, True, True)
    want = gp.llgrad(grad_X=True, grad_cov=True)
    ok = (abs(got[0] - want[0]) <= 1e-11 * abs(want[0])
          and np.allclose(got[1], want[1], rtol=1e-10, atol=1e-10 * np.abs(want[1]).max())
          and np.allclose(got[2], want[2], rtol=1e-10))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_reduction_world2_gloo():
    """N>1 host logic: shard the units over 2 ranks, all-reduce the packed partials."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_product_synthetic_matches_oracle_bitwise():
    """The product's data generator (used by bench.py) equals the golden-pinned oracle recipe."""
    from gprf_b200.synthetic import readme_dataset
    from oracle.synthetic import golden_run
    a = readme_dataset(ntrain=700, nblocks=9, ntest=100, yd=6)
    b = golden_run(700, 9, 0.1, yd=6, ntest=100)
    assert np.array_equal(a.SX, b.SX) and np.array_equal(a.SY, b.SY) and np.array_equal(a.X_obs, b.X_obs)
    assert a.neighbors == b.neighbors
    assert all(np.array_equal(x, y) for x, y in zip(a.block_idxs, b.block_idxs))
    xx = a.SX.flatten()
    assert a.x_prior(xx)[0] == b.x_prior(xx)[0]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints exactly one JSON line with the contract's keys (the CPU arm
    needs no GPU; cfg1 is BASELINE configs[0], the reference's own CPU-runnable case)."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "n=2000" in d["config"]["workload"]


def test_edge_list_counts_mutations():
    """GPRF.neighbors is a list that counts its mutations (the O(1) currency check of the device's edge
    list): every in-place edit bumps the version, reads do not, and it pickles as a plain list."""
    import pickle
    from gprf_b200.gprf import _EdgeList
    e = _EdgeList([(1, 0), (2, 1)])
    v0 = e.version
    assert e[0] == (1, 0) and len(e) == 2 and list(e) == [(1, 0), (2, 1)] and e.version == v0
    steps = [lambda: e.append((3, 2)), lambda: e.__setitem__(0, (2, 0)), lambda: e.extend([(4, 3)]),
             lambda: e.insert(1, (3, 1)), lambda: e.sort(), lambda: e.reverse(), lambda: e.pop(),
             lambda: e.remove(e[0]), lambda: e.__delitem__(0)]
    for k, f in enumerate(steps):
        f()
        assert e.version == v0 + k + 1
    e += [(5, 4)]
    assert isinstance(e, _EdgeList) and e.version == v0 + len(steps) + 1
    e.clear()
    assert len(e) == 0 and e.version == v0 + len(steps) + 2
    back = pickle.loads(pickle.dumps(_EdgeList([(1, 0)])))
    assert type(back) is list and back == [(1, 0)]


def test_traffic_figure_is_stamped_with_the_kernel_sources():
    """bench.py takes roofline.traffic from profiles/ncu_traffic.json only when that file was captured
    from the CUDA sources in the tree (hash stamp)."""
    import json
    import bench
    sha = bench.kernel_sources_sha16(bench.RESIDENT_SOURCES)
    assert len(sha) == 16 and sha == bench.kernel_sources_sha16(bench.RESIDENT_SOURCES)
    assert sha != bench.kernel_sources_sha16()               # (all headers: the other kernels' stamp)
    tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    ent = tj["cfg2"]["resident"]
    assert set(("bytes", "sources_sha16", "source", "algorithmic_bytes")) <= set(ent)
    assert ent["bytes"] < 3.5 * ent["algorithmic_bytes"]
    # the committed figure belongs to the committed sources (otherwise bench.py reports traffic: null)
    assert ent["sources_sha16"] == sha, "profiles/ncu_traffic.json is stale: re-capture (scripts/gpu_job_r02_final.sh)"
