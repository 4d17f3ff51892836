"""Stage-by-stage diagnosis of the resident unit kernel (resident.cuh) against numpy / the oracle.

    python scripts/res_diag.py [dfn wfn]

Prints the maximum error of every intermediate matrix (dumped from shared memory after each
phase), of the block exports and of the per-unit results.  Not a test: it never asserts.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gprf_b200 import GPRF, GPCov  # noqa: E402
from oracle import kernels as kern  # noqa: E402
from oracle.gprf_oracle import OracleGPRF  # noqa: E402
from oracle.kernels import GPCov as OCov  # noqa: E402


def inv_lower(L):
    return np.linalg.solve(L, np.eye(L.shape[0]))


def err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    d = np.abs(a - b)
    if not np.all(np.isfinite(a)):
        return "NON-FINITE (%d bad) " % np.sum(~np.isfinite(a))
    return "%.2e (scale %.2e)" % (d.max() if d.size else 0.0, np.abs(b).max() if b.size else 0.0)


def main():
    dfn = sys.argv[1] if len(sys.argv) > 1 else "euclidean"
    wfn = sys.argv[2] if len(sys.argv) > 2 else "se"
    rng = np.random.RandomState(1)
    sizes = [70, 45, 100, 33]
    n = sum(sizes)
    dy = 50
    if dfn == "euclidean":
        X = rng.rand(n, 2) * 0.4
        theta = dict(wfn_params=[1.2], dfn_params=[0.09, 0.12], dfn_str=dfn, wfn_str=wfn)
    else:
        X = np.column_stack([70 + 2 * rng.rand(n), 30 + 2 * rng.rand(n), 50 * rng.rand(n)])
        theta = dict(wfn_params=[1.2], dfn_params=[60.0, 40.0], dfn_str=dfn, wfn_str=wfn)
    Y = rng.randn(n, dy)
    perm = rng.permutation(n)
    blocks, o = [], 0
    for s in sizes:
        blocks.append(np.sort(perm[o:o + s]))
        o += s
    edges = [(1, 0), (2, 1), (2, 0), (3, 2)]
    nv = 0.05
    g = GPRF(X, Y, None, GPCov(**theta), nv, block_idxs=blocks, neighbors=edges, device=0)
    orc = OracleGPRF(X, Y, None, OCov(**theta), nv, block_idxs=blocks, neighbors=edges)
    cov = orc.cov
    B = len(blocks)

    def evaluate():
        out = g.llgrad(grad_X=True, grad_cov=True)
        return out, g.resident_stats()

    out, st = evaluate()
    print("resident stats (evals, fallbacks, last status):", st)
    ll0, gX0, gC0 = orc.llgrad(grad_X=True, grad_cov=True)
    print("FULL  ll", out[0], ll0, "rel", abs(out[0] - ll0) / abs(ll0))
    print("FULL  gX", err(out[1], gX0), " gC", err(out[2], gC0))

    # ---- block units -------------------------------------------------------------------------
    par = {}
    for b in range(B):
        idx = blocks[b]
        nb = len(idx)
        K = kern.kernel_matrix(X[idx], X[idx], cov) + nv * np.eye(nb)
        L = np.linalg.cholesky(K)
        W = inv_lower(L)
        Z = W @ Y[idx]
        al = W.T @ Z
        par[b] = dict(W=W, Z=Z, alpha=al, Kinv=W.T @ W)
        for ph, ref in ((2, np.tril(K)), (3, L), (4, W)):
            g.resident_debug(b, ph)
            evaluate()
            _, R2 = g.resident_dump()
            print("block %d (n=%d) phase %d R2: %s" % (b, nb, ph, err(R2[:nb, :nb], ref)))
        g.resident_debug(-1, -1)
        ex = g.resident_export(b, nb)
        print("block %d export W %s | Z %s | alpha %s | Kinv(lower) %s | logdet %.3e | q %.3e" % (
            b, err(ex["W"][:nb, :nb], W), err(ex["Z"][:nb, :dy], Z), err(ex["alpha"][:nb, :dy], al),
            err(np.tril(ex["Kinv"][:nb, :nb]), np.tril(W.T @ W)),
            abs(ex["logdet"] - 2 * np.sum(np.log(np.diag(L)))), abs(ex["q"] - np.sum(Z * Z))))
        llu, gth, gxu = g.resident_unit(b)
        l0, gx0, gc0 = orc.llgrad_unary(b, grad_X=True, grad_cov=True)
        print("block %d unit ll rel %.2e | gx %s | gth %s" % (b, abs(llu - l0) / abs(l0), err(gxu[:nb, :X.shape[1]], gx0),
                                                               err(gth[:len(gc0)], gc0)))

    # ---- pair units -----------------------------------------------------------------------------
    for e, (i, j) in enumerate(edges):
        ii, jj = blocks[i], blocks[j]
        a, b = len(ii), len(jj)
        ab8 = (a + 7) // 8 * 8
        P = par[i]
        Kji = kern.kernel_matrix(X[jj], X[ii], cov)
        Lji = Kji @ P["W"].T
        S = kern.kernel_matrix(X[jj], X[jj], cov) + nv * np.eye(b) - Lji @ Lji.T
        LS = np.linalg.cholesky(S)
        WS = inv_lower(LS)
        T = Lji @ P["W"]
        V = -WS @ T
        refs = {1: ("R1", Lji), 2: ("R2", np.tril(S)), 3: ("R2", LS), 4: ("R2", WS), 6: ("R1", T), 7: ("R1", V)}
        for ph in (1, 2, 3, 4, 6, 7):
            g.resident_debug(B + e, ph)
            evaluate()
            R1, R2 = g.resident_dump()
            which, ref = refs[ph]
            got = R1[:b, :a] if which == "R1" else R2[:b, :b]
            print("pair %d=(%d,%d) a=%d b=%d phase %d %s: %s" % (e, i, j, a, b, ph, which, err(got, ref)))
        g.resident_debug(-1, -1)
        evaluate()
        llu, gth, gxu = g.resident_unit(B + e)
        l0, gx0, gc0 = orc.llgrad_joint(i, j, grad_X=True, grad_cov=True)
        gx_loc = np.vstack([gxu[:a, :X.shape[1]], gxu[ab8:ab8 + b, :X.shape[1]]])
        print("pair %d unit ll rel %.2e | gx %s | gth %s" % (e, abs(llu - l0) / abs(l0), err(gx_loc, gx0),
                                                              err(gth[:len(gc0)], gc0)))


if __name__ == "__main__":
    main()
