"""GPU tests of the resident unit path (gprf_b200/csrc/resident.cuh): one CTA per unit out of shared
memory, pair units factoring only the Schur complement on top of block i's exported factor.

Checked against the CPU oracle (objective 1e-9 relative, gradients 1e-7 relative), against the tile
pipeline of the same library (two independent CUDA implementations of gprf.py:206-330), stage by
stage against numpy, and through every way the path hands an evaluation back to the tile pipeline
(unit too large, failed pivot -> jitter rule of gpy_linalg.py:77-97).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_oracle_props import COVS  # noqa: E402
from test_gpu_parity import build_pair, assert_parity, prod_cov  # noqa: E402


def inv_lower(L):
    return np.linalg.solve(L, np.eye(L.shape[0]))


@pytest.mark.parametrize("name", sorted(COVS))
@pytest.mark.parametrize("flags", [(True, True), (True, False), (False, True), (False, False)])
def test_resident_parity(name, flags):
    """Empty, single-point, ragged blocks, blocks beyond 16 x 8 points (second pass over the rows),
    pairs whose coupling matrix does not fit in shared memory (kept in L2 scratch); both size orders."""
    sizes = [40, 70, 0, 1, 64, 90, 128, 5, 113, 126, 150, 160, 33]
    edges = [(1, 0), (4, 1), (5, 4), (6, 5), (3, 1), (2, 1), (6, 0), (5, 3), (7, 6), (8, 5), (8, 0), (7, 2),
             (9, 6), (10, 9), (11, 10), (12, 11), (11, 4), (10, 0)]
    o, g = build_pair(name, sizes, edges, dy=50)
    kw = dict(grad_X=flags[0], grad_cov=flags[1])
    got = g.llgrad(**kw)
    ev, fb, st = g.resident_stats()
    assert ev == 1 and fb == 0 and st == 0, (ev, fb, st)
    assert_parity(o.llgrad(**kw), got, name)
    lls, jit = g.unit_results()
    for u in range(len(sizes)):
        ref = o.llgrad_unary(u)[0]
        assert abs(lls[u] - ref) <= 1e-9 * max(1.0, abs(ref))
    for e, (i, j) in enumerate(edges):
        ref = o.llgrad_joint(i, j)[0]
        assert abs(lls[len(sizes) + e] - ref) <= 1e-9 * max(1.0, abs(ref)), (e, i, j)
    # the tile pipeline (independent pdinv per unit) agrees far inside the tolerance
    g.set_resident(False)
    tile = g.llgrad(**kw)
    assert g.resident_stats()[0] == 1
    assert abs(tile[0] - got[0]) <= 1e-11 * abs(got[0])
    if flags[0]:
        assert np.abs(tile[1] - got[1]).max() <= 1e-9 * np.abs(got[1]).max()
    if flags[1]:
        assert np.abs(tile[2] - got[2]).max() <= 1e-9 * np.abs(got[2]).max()
    # deterministic: bit-identical run to run (fixed summation order, dynamic unit queue or not)
    g.set_resident(True)
    again = g.llgrad(**kw)
    assert again[0] == got[0] and np.array_equal(again[1], got[1]) and np.array_equal(again[2], got[2])


def test_resident_elementwise_gradient_error():
    """Worst element-wise relative error of gradX on entries above 1e-3 max|g| (the norm-wise
    bound of assert_parity is the north-star's tolerance; this pins the stronger statement)."""
    sizes = [100, 95, 110, 104, 99, 120]
    edges = [(1, 0), (2, 1), (3, 2), (4, 3), (5, 4), (5, 0), (3, 0)]
    o, g = build_pair("euclid_se", sizes, edges, dy=50, seed=5, nv=0.01)
    want = o.llgrad(grad_X=True, grad_cov=True)
    got = g.llgrad(grad_X=True, grad_cov=True)
    assert g.resident_stats()[1] == 0
    big = np.abs(want[1]) > 1e-3 * np.abs(want[1]).max()
    rel = np.abs(got[1] - want[1])[big] / np.abs(want[1])[big]
    assert rel.max() <= 1e-7, rel.max()
    assert np.all(np.abs(got[2] - want[2]) <= 1e-7 * np.abs(want[2]))


def test_resident_stages():
    """Shared-memory matrices after each phase, the block exports and the per-unit results."""
    from oracle import kernels as kern
    sizes = [70, 45, 100]
    edges = [(1, 0), (2, 1), (2, 0)]
    nv = 0.05
    o, g = build_pair("euclid_se", sizes, edges, dy=50, seed=2, nv=nv)
    X, Y, cov, blocks = o.X, o.Y, o.cov, o.block_idxs
    B = len(sizes)

    def run():
        g.llgrad(grad_X=True, grad_cov=True)
        assert g.resident_stats()[2] == 0

    def close(a, b, what):
        assert np.abs(a - b).max() <= 1e-10 * max(np.abs(b).max(), 1e-300), what

    par = {}
    for b in range(B):
        idx = blocks[b]
        nb = len(idx)
        K = kern.kernel_matrix(X[idx], X[idx], cov) + nv * np.eye(nb)
        L = np.linalg.cholesky(K)
        W = inv_lower(L)
        par[b] = W
        for ph, ref in ((3, L), (4, W)):
            g.resident_debug(b, ph)
            run()
            close(g.resident_dump()[1][:nb, :nb], ref, "block %d phase %d" % (b, ph))
        g.resident_debug(-1, -1)
        ex = g.resident_export(b, nb)
        Z = W @ Y[idx]
        close(ex["W"][:nb, :nb], W, "export W")
        close(ex["Z"][:nb, :50], Z, "export Z")
        close(ex["alpha"][:nb, :50], W.T @ Z, "export alpha")
        close(np.tril(ex["Kinv"][:nb, :nb]), np.tril(W.T @ W), "export Kinv")
        assert np.abs(ex["Z"][nb:]).max() == 0 and np.abs(ex["Z"][:, 50:]).max() == 0      # padding stays zero
        assert abs(ex["logdet"] - 2 * np.sum(np.log(np.diag(L)))) <= 1e-10 * abs(ex["logdet"])
        llu, gth, gxu = g.resident_unit(b)
        l0, gx0, gc0 = o.llgrad_unary(b, grad_X=True, grad_cov=True)
        assert abs(llu - l0) <= 1e-12 * abs(l0)
        close(gxu[:nb, :2], gx0, "block gx")
        close(gth[:4], gc0, "block gth")
    for e, (i, j) in enumerate(edges):
        ii, jj = blocks[i], blocks[j]
        a, b = len(ii), len(jj)
        Wi = par[i]
        Lji = kern.kernel_matrix(X[jj], X[ii], cov) @ Wi.T
        S = kern.kernel_matrix(X[jj], X[jj], cov) + nv * np.eye(b) - Lji @ Lji.T
        LS = np.linalg.cholesky(S)
        WS = inv_lower(LS)
        T = Lji @ Wi
        refs = {1: (0, Lji), 3: (1, LS), 4: (1, WS), 6: (0, T), 7: (0, -WS @ T)}
        for ph, (which, ref) in sorted(refs.items()):
            g.resident_debug(B + e, ph)
            run()
            R = g.resident_dump()[which]
            close(R[:b, :a] if which == 0 else R[:b, :b], ref, "pair %d phase %d" % (e, ph))
        g.resident_debug(-1, -1)
        run()
        llu, gth, gxu = g.resident_unit(B + e)
        l0, gx0, gc0 = o.llgrad_joint(i, j, grad_X=True, grad_cov=True)
        ab8 = (a + 7) // 8 * 8
        assert abs(llu - l0) <= 1e-12 * abs(l0)
        close(np.vstack([gxu[:a, :2], gxu[ab8:ab8 + b, :2]]), gx0, "pair gx")
        close(gth[:4], gc0, "pair gth")


def test_resident_hands_over_to_tile_pipeline():
    """A block of more than 160 points, and a unit that needs jitter: the status word sends the
    evaluation through the tile pipeline, with the reference's jitter sequence and exceptions."""
    from gprf_b200 import GPRF, LinAlgError
    from oracle.gprf_oracle import OracleGPRF
    o, g = build_pair("euclid_m32", [40, 161, 64, 90], [(1, 0), (2, 1), (3, 2)], dy=50)
    kw = dict(grad_X=True, grad_cov=True)
    assert_parity(o.llgrad(**kw), g.llgrad(**kw), "overflow")
    assert g.resident_stats() == (1, 1, 1)                       # ST_OVERFLOW
    # a pair whose two blocks fit on their own but not together in shared memory stays resident
    o, g = build_pair("euclid_se", [128, 126, 30], [(1, 0), (2, 1)], dy=50)
    assert_parity(o.llgrad(**kw), g.llgrad(**kw), "R1 in scratch")
    assert g.resident_stats() == (1, 0, 0)
    # jitter (same construction as test_jitter_sequence_and_failures)
    cov, _ = COVS["euclid_se"]
    rng = np.random.RandomState(11)
    base = rng.rand(12, 2)
    X = np.repeat(base, 8, axis=0) + 1e-9 * rng.randn(96, 2)
    Y = rng.randn(96, 5)
    blocks = [np.arange(0, 48), np.arange(48, 96)]
    s2 = cov.wfn_params[0]
    nv = -2e-4 * s2
    o = OracleGPRF(X, Y, None, cov, nv, block_idxs=blocks, neighbors=[(1, 0)])
    g = GPRF(X, Y, None, prod_cov(cov), nv, block_idxs=blocks, neighbors=[(1, 0)])
    want = o.llgrad(**kw)
    got = g.llgrad(**kw)
    assert g.resident_stats() == (1, 1, 2)                       # ST_NOTPD
    assert np.allclose(g.unit_results()[1], (s2 + nv) * 1e-3, rtol=1e-12)
    assert abs(got[0] - want[0]) <= 1e-7 * abs(want[0])
    g.noise_var = -0.5 * s2
    with pytest.raises(LinAlgError, match="even with jitter"):
        g.llgrad()
    g.noise_var = 0.05
    o.noise_var = 0.05
    assert_parity(o.llgrad(grad_X=True), g.llgrad(grad_X=True), "after failures")
    assert g.resident_stats()[2] == 0


def test_resident_reblocking_walk():
    """update_X -> on-device re-blocking -> resident evaluation, no host view of the blocks in
    between; the lazily fetched block lists equal the host partition (bit-exact membership)."""
    from gprf_b200 import GPRF, GPCov, Blocker, grid_centers
    from oracle.gprf_oracle import OracleGPRF
    from oracle.kernels import GPCov as OCov
    rng = np.random.RandomState(4)
    n, dy = 900, 50
    X = rng.rand(n, 2)
    Y = rng.randn(n, dy)
    bl = Blocker(grid_centers(9))
    th = dict(wfn_params=[1.0], dfn_params=[0.2, 0.2], dfn_str="euclidean", wfn_str="se")
    g = GPRF(X, Y, bl.block_clusters, GPCov(**th), 0.01, neighbors=bl.neighbors())
    o = OracleGPRF(X, Y, bl.block_clusters, OCov(**th), 0.01, neighbors=bl.neighbors())
    assert g._device_part == "grid"
    for step in range(4):
        Xn = np.clip(X + 0.02 * step * rng.randn(n, 2), 0, 1)
        g.update_X(Xn)
        o.update_X(Xn)
        kw = dict(grad_X=True, grad_cov=(step % 2 == 1))
        assert_parity(o.llgrad(**kw), g.llgrad(**kw), "step %d" % step)
        if step in (1, 3):
            dev = g.block_idxs
            assert len(dev) == len(o.block_idxs) and all(np.array_equal(a, b) for a, b in zip(dev, o.block_idxs))
    ev, fb, st = g.resident_stats()
    assert ev >= 4 and fb == 0


def test_resident_sharded_partial_sums():
    """Multi-GPU split on the resident path: the per-rank partial results (group sharding on
    nominal sizes) add up to the single-GPU result."""
    from gprf_b200 import GPRF
    sizes = [60, 70, 80, 90, 100, 65, 75, 85]
    edges = [(1, 0), (2, 1), (3, 2), (4, 3), (5, 4), (6, 5), (7, 6), (7, 0), (4, 0), (6, 2)]
    o, g = build_pair("euclid_se", sizes, edges, dy=50)
    kw = dict(grad_X=True, grad_cov=True)
    full = g.llgrad(**kw)
    for world in (2, 3):
        acc = [0.0, np.zeros_like(full[1]), np.zeros_like(full[2])]
        for rank in range(world):
            gs = GPRF(o.X, o.Y, None, prod_cov(o.cov), o.noise_var, block_idxs=o.block_idxs, neighbors=list(edges),
                      unit_shard=(rank, world))
            part = gs.llgrad(**kw)
            assert gs.resident_stats()[1] == 0
            acc[0] += part[0]
            acc[1] += part[1]
            acc[2] += part[2]
            gs.close()
        assert abs(acc[0] - full[0]) <= 1e-12 * abs(full[0])
        assert np.abs(acc[1] - full[1]).max() <= 1e-11 * np.abs(full[1]).max()
        assert np.abs(acc[2] - full[2]).max() <= 1e-11 * np.abs(full[2]).max()
    assert_parity(o.llgrad(**kw), full, "full")


@pytest.mark.parametrize("n,ncenters", [(2500, 36), (4200, 64)])
def test_single_cta_bucketing_equals_radix_sort_path(monkeypatch, n, ncenters):
    """Re-blocking of small problems (stable counting sort + bounds + the resident path's launch plan,
    partition.cuh: k_bucket_small) runs in ONE CTA, or from 4096 points on as a split launch (CTA 0: block
    totals + plan, four more CTAs: placement over finer sub-ranges; GPRF_BUCKET_SPLIT=0 keeps the one CTA);
    GPRF_BUCKET_SMALL=0 keeps the cub radix sort + k_block_bounds + k_res_plan launches.  Same block lists
    (bit-exact against numpy, empty blocks included), same results to the last bit, three launches fewer."""
    from gprf_b200 import GPRF, GPCov, Blocker, grid_centers
    from oracle.blocking import Blocker as OB, grid_centers as ogc
    rng = np.random.RandomState(8)
    dy = 50
    Y = rng.randn(n, dy)
    bl = Blocker(grid_centers(ncenters))
    bo = OB(np.asarray(ogc(ncenters)))
    th = dict(wfn_params=[1.0], dfn_params=[0.1, 0.1], dfn_str="euclidean", wfn_str="se")
    res = {}
    Xs = []
    for step in range(3):
        X = rng.rand(n, 2)
        if step == 1:
            X[:, 0] *= 0.6                     # leaves the right-hand blocks empty
        Xs.append(X)
    modes = (("1", "4"), ("1", "0"), ("0", "4"))
    for mode in modes:
        monkeypatch.setenv("GPRF_BUCKET_SMALL", mode[0])
        monkeypatch.setenv("GPRF_BUCKET_SPLIT", mode[1])
        g = GPRF(Xs[0], Y, bl.block_clusters, GPCov(**th), 0.01, neighbors=bl.neighbors())
        assert g._device_part == "grid"
        out = []
        for X in Xs:
            g.update_X(X)
            r = g.llgrad(grad_X=True, grad_cov=True)
            assert g.resident_stats()[1] == 0
            launches = g.last_timing()[1]
            dev = g.block_idxs
            host = bo.block_clusters(X)
            assert len(dev) == len(host) and all(np.array_equal(a, b) for a, b in zip(dev, host))
            out.append((r, launches))
        res[mode] = out
        g.close()
    for mode in modes[1:]:
        for (a, la), (b, lb) in zip(res[modes[0]], res[mode]):
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
            assert la == (lb - 4 if mode[0] == "0" else lb), (la, lb)          # 3 cub launches + bounds + plan -> 1


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_resident_random_structures(seed):
    """Randomised structures (block sizes 0..160 incl. multiples of 8 +-1, random edge sets with repeated
    parents, every covariance family by turns): resident path == oracle, and the static unit lists cover
    every unit exactly once whatever the number of CTAs they are dealt to."""
    rng = np.random.RandomState(100 + seed)
    name = sorted(COVS)[seed % len(COVS)]
    B = int(rng.randint(3, 14))
    pool = [0, 1, 7, 8, 9, 15, 16, 17, 63, 64, 65, 100, 127, 128, 129, 159, 160]
    sizes = [int(pool[rng.randint(len(pool))]) if rng.rand() < 0.5 else int(rng.randint(1, 161)) for _ in range(B)]
    pairs = [(i, j) for i in range(B) for j in range(i)]
    rng.shuffle(pairs)
    edges = pairs[:int(rng.randint(1, min(len(pairs), 3 * B) + 1))]
    o, g = build_pair(name, sizes, edges, dy=int(rng.choice([1, 7, 50, 64])), seed=seed)
    kw = dict(grad_X=True, grad_cov=True)
    got = g.llgrad(**kw)
    assert g.resident_stats() == (1, 0, 0), (sizes, edges)
    assert_parity(o.llgrad(**kw), got, "random structure %d (%s)" % (seed, name))
    lls, _ = g.unit_results()
    for e, (i, j) in enumerate(edges):
        if sizes[i] + sizes[j] == 0:
            continue
        ref = o.llgrad_joint(i, j)[0]
        assert abs(lls[B + e] - ref) <= 1e-9 * max(1.0, abs(ref)), (e, i, j, sizes[i], sizes[j])


@pytest.mark.timeout(120)
def test_two_handles_from_two_threads():
    """Two GPRF objects evaluated concurrently from two host threads (ctypes releases the GIL): resident
    launches are serialised per process - their CTAs wait for each other's results, so two grids must not
    compete for the SMs - and every evaluation returns the single-threaded result, bit for bit."""
    import threading
    o1, g1 = build_pair("euclid_se", [90, 100, 80, 110, 95, 70], [(1, 0), (2, 1), (3, 2), (4, 3), (5, 4), (5, 0)], dy=50, seed=3)
    o2, g2 = build_pair("euclid_m32", [60, 120, 100, 85], [(1, 0), (2, 1), (3, 2), (3, 0)], dy=50, seed=4)
    kw = dict(grad_X=True, grad_cov=True)
    want = [g1.llgrad(**kw), g2.llgrad(**kw)]
    assert_parity(o1.llgrad(**kw), want[0], "thread 1 reference")
    assert_parity(o2.llgrad(**kw), want[1], "thread 2 reference")
    errors = []

    def work(g, ref):
        try:
            for _ in range(60):
                got = g.llgrad(**kw)
                if not (got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])):
                    errors.append("result changed")
                    return
        except Exception as exc:        # noqa: BLE001
            errors.append(repr(exc))
    ts = [threading.Thread(target=work, args=(g1, want[0])), threading.Thread(target=work, args=(g2, want[1]))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(100)
    assert not any(t.is_alive() for t in ts), "an evaluation hangs"
    assert errors == []
    assert g1.resident_stats()[1] == 0 and g2.resident_stats()[1] == 0
