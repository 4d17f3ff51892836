"""CPU restatement of ``GPRF.llgrad`` and what it calls (oracle side).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Never imported by the
product package; the CUDA path is checked against this.

Follows gprf.py:83-375 (class, neighbours, ``update_*``, ``llgrad``,
``llgrad_unary/joint``, ``kernel``, ``dKdx``, ``dKdi``) and gprf.py:496-591
(``gaussian_llgrad``).  Two evaluation modes for one unit:

``faithful``  same loop structure as gprf.py:553-584: the derivative matrices
              are filled row by row (one kernel_deriv_wrt_xi_row call per
              (point, dim)), then the per-dimension / per-parameter
              contractions are done with dense numpy ops in the reference's
              order.
``fast``      algebraically identical, vectorised: G = A A^T - dy K^-1 and
              one elementwise contraction per derivative.
"""
from collections import defaultdict
from multiprocessing import Pool
import multiprocessing

import numpy as np

from . import kernels as kern
from .kernels import GPCov
from .linalg import pdinv, dpotrs

LOG2PI = np.log(2 * np.pi)


def symmetrize_neighbors(neighbors):
    nd = defaultdict(set)
    for (i, j) in neighbors:
        nd[i].add(j)
        nd[j].add(i)
    return nd


class OracleGPRF(object):
    def __init__(self, X, Y, block_fn, cov, noise_var, neighbor_threshold=1e-3,
                 block_idxs=None, neighbors=None, mode="fast"):
        self.X = X
        self.Y = Y
        self.block_fn = block_fn
        self.block_idxs = block_fn(X) if block_idxs is None else block_idxs
        self.n_blocks = len(self.block_idxs)
        self.cov = cov
        self.noise_var = noise_var
        self.mode = mode
        self.neighbor_threshold = neighbor_threshold
        if neighbors is not None:
            self.neighbors = neighbors
        else:
            self.compute_neighbors(neighbor_threshold)
        self.compute_neighbor_count()
        self.neighbor_dict = symmetrize_neighbors(self.neighbors)

    # -- structure (gprf.py:119-157) -----------------------------------
    def compute_neighbors(self, threshold=1e-3):
        self.neighbors = []
        if threshold == 1.0:
            return
        s2 = self.cov.wfn_params[0]
        for i in range(self.n_blocks):
            Xi = self.X[self.block_idxs[i]]
            for j in range(i):
                Kij = self.kernel(Xi, X2=self.X[self.block_idxs[j]]) / s2
                if Kij.size and np.max(np.abs(Kij)) > threshold:
                    self.neighbors.append((i, j))

    def compute_neighbor_count(self):
        cnt = defaultdict(int)
        for (i, j) in self.neighbors:
            cnt[i] += 1
            cnt[j] += 1
        self.neighbor_count = cnt

    # -- mutators (gprf.py:160-174) ------------------------------------
    def update_covs(self, covs):
        nv, sv = covs[0, :2]
        self.cov = GPCov(wfn_params=[sv], dfn_params=covs[0, 2:],
                         dfn_str=self.cov.dfn_str, wfn_str=self.cov.wfn_str)
        self.noise_var = nv

    def update_X(self, new_X, update_blocks=True, recompute_neighbors=False):
        self.X = new_X
        if self.block_fn is not None:
            self.block_idxs = self.block_fn(new_X)
        if recompute_neighbors:
            self.compute_neighbors(self.neighbor_threshold)

    # -- kernel wrappers (gprf.py:333-375) -----------------------------
    def kernel(self, X, X2=None):
        if X2 is None:
            return kern.kernel_matrix(X, X, self.cov) + np.eye(X.shape[0]) * self.noise_var
        return kern.kernel_matrix(X, X2, self.cov)

    def dKdi(self, X1, t):
        if t == 0:
            return np.eye(X1.shape[0])
        if t == 1:
            if len(self.cov.wfn_params) != 1:
                raise ValueError("gradient computation assumes a single weight-function parameter")
            return self.kernel(X1, X1) / self.cov.wfn_params[0]
        return kern.kernel_deriv_wrt_i(X1, X1, t - 2, self.cov)

    # -- one unit (gprf.py:496-591) ------------------------------------
    def gaussian_llgrad(self, X, Y, grad_X=False, grad_cov=False):
        n, dx = X.shape
        dy = Y.shape[1]
        ncov = 2 + len(self.cov.dfn_params)
        gradX = np.zeros(())
        gradC = np.zeros(())
        if n == 0:
            if grad_X:
                gradX = np.zeros(X.shape)
            if grad_cov:
                gradC = np.zeros((ncov,))
            return 0.0, gradX, gradC

        K = self.kernel(X)
        prec, L, logdet = pdinv(K)
        Alpha = dpotrs(L, Y)
        ll = -.5 * np.sum(Y * Alpha)
        ll += -.5 * dy * logdet
        ll += -.5 * dy * n * LOG2PI

        if self.mode == "faithful":
            if grad_X:
                gradX = np.zeros((n, dx))
                dK = [np.zeros(K.shape) for _ in range(dx)]
                for p in range(n):
                    for i in range(dx):
                        dK[i][p, :] = kern.kernel_deriv_wrt_xi_row(X, p, i, self.cov)
                for i in range(dx):
                    gradX[:, i] = -dy * np.sum(prec * dK[i], axis=1)
                    gradX[:, i] += np.sum(np.dot(dK[i], Alpha) * Alpha, axis=1)
            if grad_cov:
                gradC = np.zeros((ncov,))
                for t in range(ncov):
                    dKt = self.dKdi(X, t)
                    gradC[t] = .5 * np.sum(Alpha * np.dot(dKt, Alpha)) - .5 * dy * np.sum(prec * dKt)
        else:
            if grad_X or grad_cov:
                G = np.dot(Alpha, Alpha.T) - dy * prec
            if grad_X:
                gradX = np.stack([np.sum(G * dKi, axis=1)
                                  for dKi in kern.kernel_deriv_wrt_xi_rows(X, self.cov)], axis=1)
            if grad_cov:
                gradC = np.array([.5 * np.sum(G * self.dKdi(X, t)) for t in range(ncov)])
        return ll, gradX, gradC

    def llgrad_unary(self, i, **kwargs):
        idx = self.block_idxs[i]
        return self.gaussian_llgrad(self.X[idx], self.Y[idx], **kwargs)

    def llgrad_joint(self, i, j, **kwargs):
        ii, jj = self.block_idxs[i], self.block_idxs[j]
        return self.gaussian_llgrad(np.vstack([self.X[ii], self.X[jj]]),
                                    np.vstack([self.Y[ii], self.Y[jj]]), **kwargs)

    # -- the objective (gprf.py:206-296) -------------------------------
    def llgrad(self, parallel=False, local=True, **kwargs):
        kwargs.pop("sparse", None)
        if local:
            neighbors, count = self.neighbors, self.neighbor_count
        else:
            neighbors = [(i, j) for i in range(self.n_blocks) for j in range(i)]
            count = dict((i, self.n_blocks - 1) for i in range(self.n_blocks))

        if parallel:
            # gprf.py:218-233: a fresh Pool(cpu_count()) per call, unaries then pairs
            pool = Pool(processes=multiprocessing.cpu_count())
            try:
                unaries = pool.map_async(_unary_shim, [(kwargs, self, i) for i in range(self.n_blocks)]).get(9999999)
                pairs = (pool.map_async(_joint_shim, [(kwargs, self, i, j) for (i, j) in neighbors]).get(9999999)
                         if len(neighbors) > 0 else [])
                pool.close()
                pool.join()
            except KeyboardInterrupt:
                pool.terminate()
                raise
        else:
            unaries = [self.llgrad_unary(i, **kwargs) for i in range(self.n_blocks)]
            pairs = [self.llgrad_joint(i, j, **kwargs) for (i, j) in neighbors]

        ll = np.sum([p[0] for p in pairs])
        ll += np.sum([(1 - count[i]) * u[0] for (i, u) in enumerate(unaries)])

        if kwargs.get("grad_X"):
            gradX = np.zeros(self.X.shape)
            for i in range(self.n_blocks):
                gradX[self.block_idxs[i], :] -= (count[i] - 1) * unaries[i][1]
            for e, (i, j) in enumerate(neighbors):
                ni = len(self.block_idxs[i])
                gradX[self.block_idxs[i]] += pairs[e][1][:ni]
                gradX[self.block_idxs[j]] += pairs[e][1][ni:]
        else:
            gradX = np.zeros((0, 0))

        if kwargs.get("grad_cov"):
            gradCov = np.sum([p[2] for p in pairs], axis=0)
            gradCov = gradCov - np.sum([(count[i] - 1) * unaries[i][2] for i in range(self.n_blocks)], axis=0)
            gradCov = np.asarray(gradCov, dtype=float).reshape((1, -1))
        else:
            gradCov = np.zeros((0, 0))
        return ll, gradX, gradCov


def _bcm_predict(gp, kernel_fn, block_Kinvs, block_Alphas, dy, Xstar, test_noise_var, prior_kernel_fn=None):
    """Body of the closure returned by train_predictor (gprf.py:619-670): Bayesian-committee
    fusion of the per-block GP predictions of the test points' own block and its neighbours.
    ``prior_kernel_fn`` (test_cov, gprf.py:599-605,621) enters the prior covariance only; Kstar and Kss
    use the training covariance (predict_tree, gprf.py:649-654)."""
    prior_cov = (prior_kernel_fn or kernel_fn)(Xstar, Xstar)
    prior_cov = prior_cov + np.eye(prior_cov.shape[0]) * test_noise_var
    prior_prec = np.linalg.inv(prior_cov)
    prior_mean = np.zeros((Xstar.shape[0], dy))
    test_block_idxs = gp.block_fn(Xstar)
    source_blocks = set()
    for i, idxs in enumerate(test_block_idxs):
        if len(idxs) == 0:
            continue
        source_blocks.add(i)
        for j in gp.neighbor_dict[i]:
            source_blocks.add(j)
    for i in sorted(source_blocks):          # the reference iterates a set of small ints: ascending
        X = gp.X[gp.block_idxs[i]]
        Kstar = kernel_fn(Xstar, X)
        Kss = kernel_fn(Xstar, Xstar)
        if test_noise_var > 0:
            Kss = Kss + np.eye(Kss.shape[0]) * gp.noise_var       # gprf.py:653-655 adds the MODEL's noise
        mean = np.dot(Kstar, block_Alphas[i])
        cov = Kss - np.dot(Kstar, np.dot(block_Kinvs[i], Kstar.T))
        prec = np.linalg.inv(cov)
        pp = np.linalg.inv(Kss)
        prior_mean += np.dot(prec, mean)
        prior_prec += prec - pp
    final_cov = np.linalg.inv(prior_prec)
    return np.dot(final_cov, prior_mean), final_cov


def _train_predictor(self, test_cov=None, Y=None):
    """gprf.py:593-672.  Per block: Kinv = inv(k(X_b) + nv I), Alpha = Kinv Y_b; the returned
    ``predict(Xstar, test_noise_var=0.0, local=False)`` gives (mean, cov) of the BCM fusion.
    (As shipped the reference passes ``block=`` to ``kernel``, which has no such parameter -
    SURVEY.md section 8c; this is the evident intent.)"""
    Y = self.Y if Y is None else Y
    tcov = self.cov if test_cov is None else test_cov
    block_Kinvs, block_Alphas = [], []
    for idxs in self.block_idxs:
        K = self.kernel(self.X[idxs])
        Kinv = np.linalg.inv(K) if len(idxs) else np.zeros((0, 0))
        block_Kinvs.append(Kinv)
        block_Alphas.append(np.dot(Kinv, Y[idxs]))

    def kernel_fn(A, B):
        return kern.kernel_matrix(A, B, self.cov)

    def prior_kernel_fn(A, B):
        return kern.kernel_matrix(A, B, tcov)

    def predict(Xstar, test_noise_var=0.0, local=False):
        return _bcm_predict(self, kernel_fn, block_Kinvs, block_Alphas, Y.shape[1], Xstar, test_noise_var,
                            prior_kernel_fn=prior_kernel_fn)
    return predict


OracleGPRF.train_predictor = _train_predictor


def _unary_shim(arg):
    return OracleGPRF.llgrad_unary(*arg[1:], **arg[0])


def _joint_shim(arg):
    return OracleGPRF.llgrad_joint(*arg[1:], **arg[0])
