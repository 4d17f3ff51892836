"""Properties of the CPU oracle that the reference offers no golden values for."""
import numpy as np
import pytest
from numpy.linalg import LinAlgError

from oracle import kernels as kern
from oracle.kernels import GPCov
from oracle.blocking import Blocker, grid_centers, pdtree_cluster, cluster_rpc
from oracle.gprf_oracle import OracleGPRF
from oracle.linalg import jitchol

COVS = {
    "euclid_se": (GPCov([1.3], [0.21, 0.17], "euclidean", "se"), 2),
    "euclid_m32": (GPCov([0.8], [0.25, 0.31], "euclidean", "matern32"), 2),
    "lld_m32": (GPCov([1.0], [40.0, 25.0], "lld", "matern32"), 3),
    "lld_se": (GPCov([1.1], [55.0, 30.0], "lld", "se"), 3),
}


def make_points(n, dx, rng):
    if dx == 2:
        return rng.rand(n, 2)
    return np.column_stack([80 + 0.8 * rng.rand(n), 35 + 0.6 * rng.rand(n), 40 * rng.rand(n)])


def small_gprf(name, n=60, nb=4, dy=5, seed=3, mode="fast"):
    cov, dx = COVS[name]
    rng = np.random.RandomState(seed)
    X = make_points(n, dx, rng)
    Y = rng.randn(n, dy)
    owner = rng.randint(0, nb, size=n)
    blocks = [np.arange(n)[owner == b] for b in range(nb)]
    nbrs = [(i, j) for i in range(nb) for j in range(i) if (i + j) % 2 == 1]
    return OracleGPRF(X, Y, None, cov, 0.05, block_idxs=blocks, neighbors=nbrs, mode=mode)


def test_great_circle_doctests():
    # run_seismic.py:24-33
    deg = lambda a, b: np.degrees(kern.great_circle_km(a[0], a[1], b[0], b[1]) / kern.EARTH_RADIUS_KM)
    assert int(deg((10, 0), (20, 0)) + 1e-9) == 10
    assert int(deg((10, 0), (10, 45)) + 1e-9) == 45
    assert int(deg((-78, -12), (-10.25, 52))) == 86
    assert deg((132.86521, -0.45606493), (132.86521, -0.45606493)) < 1e-4
    assert deg((127.20443, 2.8123965), (127.20443, 2.8123965)) < 1e-4


@pytest.mark.parametrize("m,edges", [(2, 6), (3, 20), (5, 72), (10, 342), (20, 1482)])
def test_grid_edge_counts(m, edges):
    b = Blocker(np.asarray(grid_centers(m * m)))
    nb = b.neighbors(True)
    assert len(nb) == edges
    assert all(j < i for (i, j) in nb)


@pytest.mark.parametrize("name", sorted(COVS))
def test_faithful_equals_fast(name):
    a = small_gprf(name, mode="faithful").llgrad(grad_X=True, grad_cov=True)
    b = small_gprf(name, mode="fast").llgrad(grad_X=True, grad_cov=True)
    assert a[0] == b[0]
    np.testing.assert_allclose(a[1], b[1], rtol=1e-9, atol=1e-9 * np.abs(a[1]).max())
    np.testing.assert_allclose(a[2], b[2], rtol=1e-9)
    assert a[2].shape == (1, 4)


@pytest.mark.parametrize("name", sorted(COVS))
def test_gradients_by_finite_differences(name):
    gp = small_gprf(name)
    ll, gX, gC = gp.llgrad(grad_X=True, grad_cov=True)
    X0 = gp.X.copy()
    rng = np.random.RandomState(0)
    for _ in range(6):
        p, i = rng.randint(X0.shape[0]), rng.randint(X0.shape[1])
        h = 1e-6 * max(1.0, abs(X0[p, i]))
        Xp, Xm = X0.copy(), X0.copy()
        Xp[p, i] += h
        Xm[p, i] -= h
        gp.X = Xp
        fp = gp.llgrad()[0]
        gp.X = Xm
        fm = gp.llgrad()[0]
        fd = (fp - fm) / (2 * h)
        assert abs(fd - gX[p, i]) <= 2e-5 * max(1.0, abs(gX[p, i]), np.abs(gX).max() * 1e-2)
    gp.X = X0
    th0 = np.concatenate([[gp.noise_var], gp.cov.wfn_params, gp.cov.dfn_params]).reshape(1, -1)
    for t in range(4):
        h = 1e-6 * th0[0, t]
        tp, tm = th0.copy(), th0.copy()
        tp[0, t] += h
        tm[0, t] -= h
        gp.update_covs(tp)
        fp = gp.llgrad()[0]
        gp.update_covs(tm)
        fm = gp.llgrad()[0]
        fd = (fp - fm) / (2 * h)
        assert abs(fd - gC[0, t]) <= 1e-5 * max(1.0, abs(gC[0, t]))


def test_empty_block_and_shapes():
    gp = small_gprf("euclid_se")
    gp.block_idxs[1] = gp.block_idxs[1][:0]
    ll, gX, gC = gp.llgrad(grad_X=True, grad_cov=True)
    assert np.isfinite(ll) and gX.shape == gp.X.shape and gC.shape == (1, 4)
    ll2, gX2, gC2 = gp.llgrad()
    assert ll2 == ll and gX2.shape == (0, 0) and gC2.shape == (0, 0)


def test_nonlocal_uses_all_pairs():
    gp = small_gprf("euclid_se")
    full = gp.llgrad(local=False)[0]
    gp2 = small_gprf("euclid_se")
    gp2.neighbors = [(i, j) for i in range(4) for j in range(i)]
    gp2.compute_neighbor_count()
    assert full == gp2.llgrad()[0]


def test_jitchol_sequence():
    rng = np.random.RandomState(1)
    B = rng.randn(30, 5)
    A = B.dot(B.T)                                  # rank 5, PSD
    A = A - 2e-4 * np.mean(np.diag(A)) * np.eye(30)  # min eig ~ -2e-4*mean(diag)
    L, jit = jitchol(A, return_jitter=True)
    assert np.isclose(jit, np.mean(np.diag(A)) * 1e-3)   # 1e-6,1e-5,1e-4 fail; 1e-3 succeeds
    np.testing.assert_allclose(L.dot(L.T), A + jit * np.eye(30), atol=1e-10)
    with pytest.raises(LinAlgError):
        jitchol(A - 0.5 * np.mean(np.diag(A)) * np.eye(30))
    with pytest.raises(LinAlgError):
        jitchol(A - 2 * np.max(np.diag(A)) * np.eye(30))


def test_pdtree_and_rpc_partition():
    rng = np.random.RandomState(5)
    X = np.column_stack([rng.uniform(-30, 330, 500), rng.uniform(-60, 60, 500), rng.rand(500) * 100])
    idxs, reblock = pdtree_cluster(X, blocksize=60)
    assert sorted(np.concatenate(idxs).tolist()) == list(range(500))
    assert all(30 <= len(i) < 60 for i in idxs)
    X_before = X.copy()
    again = reblock(X)
    assert np.array_equal(X, X_before)
    assert all(np.array_equal(a, b) for a, b in zip(idxs, again))
    np.random.seed(2)
    parts, splits = cluster_rpc(X[:, :2], np.arange(500), target_size=70)
    assert sorted(np.concatenate(parts).tolist()) == list(range(500))
    parts2, _ = cluster_rpc(X[:, :2], np.arange(500), target_size=70, fixed_split=splits)
    assert all(np.array_equal(a, b) for a, b in zip(parts, parts2))
