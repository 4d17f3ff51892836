"""Synthetic data and priors of the reference's experiment driver (oracle side).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Follows synthetic.py:103-114,139-153 (``sample_y`` dense branch,
``sample_synthetic`` uniform branch) and gprfopt.py:19-74,172-182
(``SampledData``: train/test split, X_obs, centres, x_prior).  The legacy
``np.random.seed`` global stream is used on purpose - the golden objective
values in ``tests/golden/`` are reproducible only with it.
"""
import numpy as np

from .kernels import GPCov, kernel_matrix
from .linalg import jitchol
from .blocking import Blocker, grid_centers
from .gprf_oracle import OracleGPRF


def sample_y(X, cov, noise_var, yd):
    """Y = chol(K + nv I) Z (synthetic.py:103-114; dense branch, n < 40000)."""
    K = kernel_matrix(X, X, cov)
    K[np.diag_indices_from(K)] += noise_var
    L = jitchol(K)
    del K
    Z = np.random.randn(X.shape[0], yd)
    return np.dot(L, Z)


def sample_synthetic(seed=1, n=400, xd=2, yd=10, lscale=0.1, noise_var=0.01):
    """Uniform X in the unit square + GP-prior Y (synthetic.py:139-153, seed < 1000)."""
    if seed >= 1000:
        raise NotImplementedError("shaped datasets (seed >= 1000) are outside the hot-path scope")
    np.random.seed(seed)
    X = np.random.rand(n, xd)
    cov = GPCov(wfn_params=[1.0], dfn_params=[lscale, lscale], dfn_str="euclidean", wfn_str="se")
    return X, sample_y(X, cov, noise_var, yd), cov


class SampledData(object):
    """gprfopt.py:19-74,172-182."""

    def __init__(self, noise_var=0.01, n=30, ntrain=20, lscale=0.5, obs_std=0.05, yd=10, seed=1):
        self.noise_var, self.n, self.ntrain, self.lscale = noise_var, n, ntrain, lscale
        Xfull, Yfull, cov = sample_synthetic(n=n, noise_var=noise_var, yd=yd, lscale=lscale, seed=seed)
        self.cov = cov
        self.SX, self.SY = Xfull[:ntrain, :], Yfull[:ntrain, :]
        self.Xtest, self.Ytest = Xfull[ntrain:, :], Yfull[ntrain:, :]
        self.block_idxs = None
        self.obs_std = obs_std
        np.random.seed(seed)
        self.X_obs = self.SX + np.random.randn(*self.SX.shape) * obs_std

    def set_centers(self, centers):
        self.centers = np.asarray(centers)
        b = Blocker(self.centers)
        self.block_idxs = b.block_clusters(self.X_obs)
        self.reblock = b.block_clusters
        self.neighbors = b.neighbors(diag_connections=True)

    def build_gprf(self, X=None, local_dist=1e-4, cls=OracleGPRF, **extra):
        """gprfopt.py:55-74; ``cls`` lets tests build the CUDA-backed GPRF the same way."""
        if X is None:
            X = self.X_obs
        return cls(X, self.SY, self.reblock, self.cov, self.noise_var,
                   neighbor_threshold=local_dist, block_idxs=self.block_idxs,
                   neighbors=self.neighbors if local_dist < 1.0 else [], **extra)

    def prediction_error(self, X=None, cov=None, local_dist=1.0):
        """gprfopt.py:121-170 restated: (smse, msll_block, msll_block_diag)."""
        import scipy.stats
        assert cov is None, "the oracle restatement covers the task=x use (cov fixed at its true value)"
        gprf = self.build_gprf(X=X, local_dist=local_dist)
        p = gprf.train_predictor()
        test_blocks = self.reblock(self.Xtest)

        def gaussian_ll(Y, M, Cm):
            ntest, yd = Y.shape
            R = Y - M
            ll = -.5 * np.sum(np.linalg.inv(Cm) * np.dot(R, R.T))
            ll -= .5 * yd * np.linalg.slogdet(Cm)[1]
            return ll - .5 * yd * ntest * np.log(2 * np.pi)

        ll_block = ll_block_diag = se_block = 0.0
        for idxs in test_blocks:
            if len(idxs) == 0:
                continue
            Xt, Yt = self.Xtest[idxs], self.Ytest[idxs]
            PM, PC = p(Xt, test_noise_var=self.noise_var)
            ll_block += gaussian_ll(Yt, PM, PC)
            ll_block_diag += gaussian_ll(Yt, PM, np.diag(np.diag(PC)))
            se_block += np.sum((Yt - PM) ** 2)
        ntest, yd = self.Ytest.shape
        Ymean, Ystd = np.mean(self.SY, axis=0), np.std(self.SY, axis=0)
        smse = se_block / np.sum((self.Ytest - Ymean) ** 2)
        mll_baseline = sum(np.sum(scipy.stats.norm(loc=Ymean[i], scale=Ystd[i]).logpdf(self.Ytest[:, i]))
                           for i in range(yd)) / (ntest * yd)
        return smse, ll_block / (ntest * yd) - mll_baseline, ll_block_diag / (ntest * yd) - mll_baseline

    def x_prior(self, xx):
        flat = self.X_obs.flatten()
        r = (xx - flat) / self.obs_std
        ll = -.5 * np.sum(r ** 2) - .5 * len(xx) * np.log(2 * np.pi * self.obs_std ** 2)
        return ll, -(xx - flat) / self.obs_std ** 2


def golden_run(ntrain, nblocks, local_dist, seed=0, yd=50, ntest=500, noise_var=0.01):
    """The data set of one ``*_gprf0`` golden run (gprfopt_analyze.py:206-207: lscale
    = 6/sqrt(n), obs_std = 2/sqrt(n))."""
    lscale = 6.0 / np.sqrt(ntrain)
    obs_std = 2.0 / np.sqrt(ntrain)
    sd = SampledData(noise_var=noise_var, n=ntrain + ntest, ntrain=ntrain, lscale=lscale,
                     obs_std=obs_std, yd=yd, seed=seed)
    sd.set_centers(grid_centers(nblocks))
    return sd
