"""Synthetic experiment data (the reference's ``synthetic.sample_synthetic`` and
``gprfopt.SampledData``; synthetic.py:103-114,139-153, gprfopt.py:19-74,172-182).

Host-side set-up code for drivers, benchmarks and examples - not part of the
per-evaluation hot path.  The legacy global ``np.random.seed`` stream is used
because the reference's shipped results are keyed to it.
"""
import numpy as np
from scipy.linalg import lapack

from .blocking import Blocker, grid_centers
from .cov import GPCov
from .gprf import GPRF


def _se_gram(X, lscales, signal_var):
    r2 = np.zeros((X.shape[0], X.shape[0]))
    for i in range(X.shape[1]):
        d = (X[:, i][:, None] - X[:, i][None, :]) / lscales[i]
        d *= d
        r2 += d
    np.sqrt(r2, out=r2)          # mirror r = sqrt(r2); w = exp(-r*r) of the evaluator
    r2 *= r2
    np.negative(r2, out=r2)
    np.exp(r2, out=r2)
    r2 *= signal_var
    return r2


def sample_y_device(X, cov, noise_var, yd, device=0, Z=None):
    """GP-prior draw Y = jitchol(K + nv I) Z on the GPU, any covariance family, any n whose n x n
    matrix fits in HBM (8 n^2 bytes twice: n = 20500 needs 7 GB, n = 80500 about 105 GB of the 180).

    One unit holding every point runs through the batched Cholesky of the evaluation path (kernel
    matrix generated in the factorisation's epilogues, jitter rule of gpy_linalg.py:77-97), then
    ``gprf_unit_lmul`` applies L to the draw.  This is the dense branch of the reference's sample_y
    (synthetic.py:106-114) for n < 40000 AND the stated replacement of its CHOLMOD branch
    (synthetic.py:115-135, n >= 40000): the sparse factorisation there is an approximation of this
    exact draw (kernel entries beyond 4 lengthscales dropped) whose samples depend on CHOLMOD's
    fill-reducing permutation, so the golden runs with n >= 40000 cannot be reproduced by anyone
    without that library; the exact draw can.
    """
    X = np.ascontiguousarray(X, dtype=np.float64)
    n = X.shape[0]
    if Z is None:
        Z = np.random.randn(n, yd)
    g = GPRF(X, np.ascontiguousarray(Z), None, cov, noise_var, block_idxs=[np.arange(n)], neighbors=[],
             device=device, device_blocks=False)
    try:
        g.set_resident(False)
        g.llgrad()
        Y = np.empty((n, yd), dtype=np.float64)
        g._check(g._lib.gprf_unit_lmul(g._h, 0, Y.ctypes.data))
    finally:
        g.close()
    return Y


def sample_y_local(X, cov, noise_var, yd, block_idxs, device=0, Z=None):
    """Draw from the block-local model: independently per block, Y_b = jitchol(K_bb + nv I) Z_b
    (the parity data of the n = 200000 shape, where an exact draw of the full GP
    is out of reach).  All blocks are factored by ONE device evaluation."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n = X.shape[0]
    if Z is None:
        Z = np.random.randn(n, yd)
    g = GPRF(X, np.ascontiguousarray(Z), None, cov, noise_var, block_idxs=block_idxs, neighbors=[], device=device,
             device_blocks=False)
    Y = np.zeros((n, yd), dtype=np.float64)
    try:
        g.set_resident(False)
        g.llgrad()
        for b, idx in enumerate(block_idxs):
            if len(idx) == 0:
                continue
            out = np.empty((len(idx), yd), dtype=np.float64)
            g._check(g._lib.gprf_unit_lmul(g._h, b, out.ctypes.data))
            Y[idx] = out
    finally:
        g.close()
    return Y


def sample_y(X, cov, noise_var, yd, device=None):
    """GP-prior draw Y = chol(K + nv I) Z with Z = randn(n, yd) (synthetic.py:103-137).
    ``device`` = CUDA ordinal: on the GPU (sample_y_device).  ``device=None``: host numpy, as the
    reference's dense branch does it - the synthetic (euclidean, se) family only; this is what the
    CPU-only legs (bench.py --impl reference, the oracle tests) use to build their input data."""
    if device is not None:
        return sample_y_device(X, cov, noise_var, yd, device=device)
    if cov.dfn_str != "euclidean" or cov.wfn_str != "se":
        raise NotImplementedError("host sampling covers the synthetic (euclidean, se) family; pass device=")
    K = _se_gram(np.asarray(X, dtype=float), cov.dfn_params, cov.wfn_params[0])
    K[np.diag_indices_from(K)] += noise_var
    L, info = lapack.dpotrf(K, lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("prior covariance is not positive definite")
    Z = np.random.randn(X.shape[0], yd)
    return np.dot(L, Z)


def sample_synthetic(seed=1, n=400, xd=2, yd=10, lscale=0.1, noise_var=0.01, device=None):
    if seed >= 1000:
        raise NotImplementedError("shaped synthetic sets (seed >= 1000) are not provided")
    np.random.seed(seed)
    X = np.random.rand(n, xd)
    cov = GPCov(wfn_params=[1.0], dfn_params=[lscale, lscale], dfn_str="euclidean", wfn_str="se")
    return X, sample_y(X, cov, noise_var, yd, device=device), cov


class SampledData(object):
    """Train/test split, noisy observed locations, grid blocks and the location prior."""

    def __init__(self, noise_var=0.01, n=30, ntrain=20, lscale=0.5, obs_std=0.05, yd=10, seed=1, device=None):
        self.noise_var, self.n, self.ntrain, self.lscale, self.obs_std = noise_var, n, ntrain, lscale, obs_std
        Xfull, Yfull, self.cov = sample_synthetic(n=n, noise_var=noise_var, yd=yd, lscale=lscale, seed=seed,
                                                  device=device)
        self.SX, self.SY = Xfull[:ntrain, :], Yfull[:ntrain, :]
        self.Xtest, self.Ytest = Xfull[ntrain:, :], Yfull[ntrain:, :]
        np.random.seed(seed)
        self.X_obs = self.SX + np.random.randn(*self.SX.shape) * obs_std
        self.block_idxs = None
        self.neighbors = None

    def set_centers(self, centers):
        self.centers = np.asarray(centers)
        blocker = Blocker(self.centers)
        self.block_idxs = blocker.block_clusters(self.X_obs)
        self.reblock = blocker.block_clusters
        self.neighbors = blocker.neighbors(diag_connections=True)

    def build_gprf(self, X=None, cov=None, local_dist=1e-4, cls=GPRF, **extra):
        X = self.X_obs if X is None else X
        noise_var = self.noise_var
        if cov is None:
            cov = self.cov
        else:
            cov = np.asarray(cov)
            noise_var = cov[0, 0]
            cov = GPCov(wfn_params=[cov[0, 1]], dfn_params=cov[0, 2:], dfn_str="euclidean", wfn_str="se")
        return cls(X, self.SY, self.reblock, cov, noise_var, neighbor_threshold=local_dist,
                   block_idxs=self.block_idxs, neighbors=self.neighbors if local_dist < 1.0 else [], **extra)


    def prediction_error(self, X=None, cov=None, local_dist=1.0, cls=GPRF, **extra):
        """gprfopt.py:121-170: (smse, msll_block, msll_block_diag) of the BCM predictions on the
        held-out points, block by block of the test set."""
        import scipy.stats
        gprf = self.build_gprf(X=X, cov=cov, local_dist=local_dist, cls=cls, **extra)
        p = gprf.train_predictor()
        test_blocks = self.reblock(self.Xtest)

        def gaussian_ll(Y, M, Cm):
            ntest, yd = Y.shape
            P = np.linalg.inv(Cm)
            R = Y - M
            ll = -.5 * np.sum(P * np.dot(R, R.T))
            ll -= .5 * yd * np.linalg.slogdet(Cm)[1]
            ll -= .5 * yd * ntest * np.log(2 * np.pi)
            return ll

        ll_block = ll_block_diag = se_block = 0.0
        for idxs in test_blocks:
            if len(idxs) == 0:              # (the reference would invert a 0x0 matrix here: harmless no-op)
                continue
            Xt, Yt = self.Xtest[idxs], self.Ytest[idxs]
            PM, PC = p(Xt, test_noise_var=self.noise_var)
            ll_block += gaussian_ll(Yt, PM, PC)
            ll_block_diag += gaussian_ll(Yt, PM, np.diag(np.diag(PC)))
            se_block += np.sum((Yt - PM) ** 2)
        ntest, yd = self.Ytest.shape
        Ymean = np.mean(self.SY, axis=0)
        smse = se_block / np.sum((self.Ytest - Ymean) ** 2)
        Ystd = np.std(self.SY, axis=0)
        ll_baseline = np.sum([np.sum(scipy.stats.norm(loc=Ymean[i], scale=Ystd[i]).logpdf(self.Ytest[:, i]))
                              for i in range(yd)])
        mll_baseline = ll_baseline / (ntest * yd)
        if hasattr(gprf, "close"):
            gprf.close()
        return smse, ll_block / (ntest * yd) - mll_baseline, ll_block_diag / (ntest * yd) - mll_baseline

    def x_prior(self, xx):
        resid = xx - self.X_obs.ravel()
        ll = -.5 * np.sum((resid / self.obs_std) ** 2) - .5 * len(xx) * np.log(2 * np.pi * self.obs_std ** 2)
        return ll, -resid / self.obs_std ** 2


def readme_dataset(ntrain=10000, nblocks=100, ntest=500, yd=50, seed=0, noise_var=0.01, device=None):
    """The reference's README / results configuration: lscale = 6/sqrt(n), obs_std = 2/sqrt(n)
    (gprfopt_analyze.py:206-207), grid blocks, 8-connected edges."""
    sd = SampledData(noise_var=noise_var, n=ntrain + ntest, ntrain=ntrain, lscale=6.0 / np.sqrt(ntrain),
                     obs_std=2.0 / np.sqrt(ntrain), yd=yd, seed=seed, device=device)
    sd.set_centers(grid_centers(nblocks))
    return sd
