"""Algebra of the resident (shared-memory) unit kernel, checked on the CPU against the oracle.

``gprf_b200/csrc/resident.cuh`` evaluates a pair unit [block i ; block j] (gprf.py:310-330) from
block i's own factorisation plus the Schur complement of block j - the formulas below, in the same
order.  This file restates them with dense numpy operations (no CUDA involved) and compares the
result with ``oracle.gprf_oracle.OracleGPRF.gaussian_llgrad`` on the stacked pair, so that a parity
failure of the CUDA kernel can be told apart from a mistake in the algebra it implements.
"""
import numpy as np
import pytest

from oracle import kernels as kern
from oracle.gprf_oracle import OracleGPRF
from oracle.kernels import GPCov

LOG2PI = np.log(2 * np.pi)


def inverse_lower_inplace(L):
    """W = L^-1 by descending columns, overwriting L (the order the kernel uses)."""
    W = L.copy()
    n = W.shape[0]
    for k in range(n - 1, -1, -1):
        wkk = 1.0 / W[k, k]
        col = -(W[k + 1:, k + 1:] @ W[k + 1:, k]) * wkk     # uses L[k+1:, k] and W of the trailing square
        W[k + 1:, k] = col
        W[k, k] = wkk
    return W


def block_parent(Xi, Yi, cov, nv):
    a = Xi.shape[0]
    K = kern.kernel_matrix(Xi, Xi, cov) + nv * np.eye(a)
    L = np.linalg.cholesky(K)
    W = inverse_lower_inplace(L)
    Z = W @ Yi
    return dict(W=W, Z=Z, alpha=W.T @ Z, Kinv=W.T @ W, logdet=2 * np.sum(np.log(np.diag(L))), q=np.sum(Z * Z))


def grad_from_G(G, X, cov):
    """gradX / grad theta from the LOWER triangle of G only: row sums + column sums."""
    n, dx = X.shape
    ncov = 2 + len(cov.dfn_params)
    gX = np.zeros((n, dx))
    dK = kern.kernel_deriv_wrt_xi_rows(X, cov)          # dK[d][p, q] = dk(x_p, x_q)/dx_{p, d}
    low = np.tril(np.ones((n, n)), -1)
    for d in range(dx):
        T = G * dK[d] * low                              # strictly lower entries
        gX[:, d] += T.sum(axis=1)                        # rows: point p
        gX[:, d] -= T.sum(axis=0)                        # columns: dk/dx_q = -dk/dx_p for both families?
    return gX


def pair_schur(Xi, Yi, Xj, Yj, cov, nv):
    a, b = Xi.shape[0], Xj.shape[0]
    dy = Yi.shape[1]
    P = block_parent(Xi, Yi, cov, nv)
    Kji = kern.kernel_matrix(Xj, Xi, cov)
    Lji = Kji @ P["W"].T                                 # P1
    S = kern.kernel_matrix(Xj, Xj, cov) + nv * np.eye(b) - Lji @ Lji.T      # P2
    LS = np.linalg.cholesky(S)                           # P3
    WS = inverse_lower_inplace(LS)
    R = Yj - Lji @ P["Z"]                                # P4
    Zj = WS @ R
    q = P["q"] + np.sum(Zj * Zj)
    logdet = P["logdet"] + 2 * np.sum(np.log(np.diag(LS)))
    ll = -.5 * q - .5 * dy * logdet - .5 * dy * (a + b) * LOG2PI
    T = Lji @ P["W"]                                     # P6
    V = -WS @ T
    alpha_j = WS.T @ Zj                                  # P7
    alpha_i = P["alpha"] + V.T @ Zj
    alpha = np.vstack([alpha_i, alpha_j])
    Kinv = np.zeros((a + b, a + b))                      # P8
    Kinv[:a, :a] = P["Kinv"] + V.T @ V
    Kinv[a:, :a] = WS.T @ V
    Kinv[:a, a:] = Kinv[a:, :a].T
    Kinv[a:, a:] = WS.T @ WS
    G = alpha @ alpha.T - dy * Kinv
    return ll, alpha, Kinv, G


@pytest.mark.parametrize("dfn,wfn", [("euclidean", "se"), ("euclidean", "matern32"), ("lld", "matern32")])
def test_pair_schur_matches_oracle(dfn, wfn):
    rng = np.random.RandomState(3)
    a, b, dy = 37, 29, 6
    if dfn == "euclidean":
        X = rng.rand(a + b, 2) * 0.3
        cov = GPCov([1.3], [0.11, 0.07], dfn, wfn)
    else:
        X = np.column_stack([70 + rng.rand(a + b), 30 + rng.rand(a + b), 40 * rng.rand(a + b)])
        cov = GPCov([1.3], [45.0, 30.0], dfn, wfn)
    Y = rng.randn(a + b, dy)
    nv = 0.05
    ll, alpha, Kinv, G = pair_schur(X[:a], Y[:a], X[a:], Y[a:], cov, nv)
    orc = OracleGPRF(X, Y, None, cov, nv, block_idxs=[np.arange(a + b)], neighbors=[])
    ll0, gX0, gC0 = orc.gaussian_llgrad(X, Y, grad_X=True, grad_cov=True)
    assert abs(ll - ll0) <= 1e-10 * abs(ll0)
    K = orc.kernel(X)
    assert np.allclose(Kinv, np.linalg.inv(K), rtol=0, atol=1e-8 * np.abs(Kinv).max())
    assert np.allclose(alpha, np.linalg.solve(K, Y), rtol=0, atol=1e-8 * np.abs(alpha).max())
    # gradients from the lower triangle of G: row sums + column sums with dk/dx_q evaluated as such
    n = a + b
    dK = kern.kernel_deriv_wrt_xi_rows(X, cov)
    gX = np.zeros_like(X)
    low = np.tril(np.ones((n, n)), -1)
    for d in range(X.shape[1]):
        gX[:, d] = (G * dK[d] * low).sum(axis=1) + (G * dK[d].T * low).sum(axis=0)
    assert np.allclose(gX, gX0, rtol=0, atol=1e-9 * np.abs(gX0).max())
    gC = np.zeros(2 + len(cov.dfn_params))
    gC[0] = .5 * np.trace(G)
    Kn = K - nv * np.eye(n)
    gC[1] = (.5 * np.sum(np.diag(G) * np.diag(Kn)) + np.sum(G * Kn * low)) / cov.wfn_params[0]
    for t in range(len(cov.dfn_params)):
        gC[2 + t] = np.sum(G * kern.kernel_deriv_wrt_i(X, X, t, cov) * low)
    assert np.allclose(gC, gC0, rtol=1e-9, atol=1e-9 * np.abs(gC0).max())


def test_inverse_lower_inplace():
    rng = np.random.RandomState(0)
    A = rng.randn(23, 23)
    L = np.linalg.cholesky(A @ A.T + 23 * np.eye(23))
    assert np.allclose(inverse_lower_inplace(L) @ L, np.eye(23), atol=1e-12)
