"""Decide whether a ``block_fn`` can be evaluated on the GPU bit-exactly.

The reference decides block membership with numpy expressions whose only
implementation-defined step is the BLAS dot product over the dx (= 2 or 3)
coordinates (``np.dot`` inside pair_distances, block_clustering.py:4-5, and
inside PDTree.recluster, pdtree_clustering.py:70).  This module recognises this
package's partitioners, extracts their parameters and *probes* how the host
BLAS rounds that short dot product, by comparing numpy's own results on the
actual data against candidates computed in exact rational arithmetic:

    mode 0   acc = x0*c0; acc = fma(x1, c1, acc); ...        (k-ascending FMA)
    mode 1   acc = x_{d-1}*c_{d-1}; acc = fma(x_{d-2}, ...)   (k-descending FMA)
    mode 2   every product and sum rounded separately         (no FMA)

If no candidate reproduces numpy on every probed entry the function returns
None and block assignment stays on the host.  (The GPRF constructor
additionally checks the complete device partition against the host partition
of the construction-time X.)
"""
from fractions import Fraction

import numpy as np

from .blocking import Blocker, PDTree

N_PROBE = 48


def _cand(x, c, mode):
    """Dot product of two short float vectors under rounding model ``mode`` (exactly rounded)."""
    d = len(x)
    if mode == 2:
        acc = float(x[0]) * float(c[0])
        for i in range(1, d):
            acc = acc + float(x[i]) * float(c[i])
        return acc
    order = range(d) if mode == 0 else range(d - 1, -1, -1)
    order = list(order)
    acc = float(x[order[0]]) * float(c[order[0]])
    for i in order[1:]:
        acc = float(Fraction(float(x[i])) * Fraction(float(c[i])) + Fraction(acc))   # one rounding = fma
    return acc


def _pick_mode(rows, cols, got):
    """rows: (m, d), cols: (k, d), got[m, k] = numpy's dot products.  First mode that matches all."""
    for mode in (0, 1, 2):
        ok = True
        for a in range(rows.shape[0]):
            for b in range(cols.shape[0]):
                if _cand(rows[a], cols[b], mode) != got[a, b]:
                    ok = False
                    break
            if not ok:
                break
        if ok:
            return mode
    return None


def _sample(n, k):
    if n <= k:
        return np.arange(n)
    return np.unique(np.linspace(0, n - 1, k).astype(np.int64))


def _describe_grid(blocker, X):
    C = np.ascontiguousarray(blocker.block_centers, dtype=np.float64)
    if C.ndim != 2 or C.shape[1] != X.shape[1] or X.shape[1] > 3:
        return None
    full = np.dot(X, C.T)                       # the reference's own call, same shapes
    ri, ci = _sample(X.shape[0], N_PROBE), _sample(C.shape[0], N_PROBE)
    mode = _pick_mode(X[ri], C[ci], full[np.ix_(ri, ci)])
    if mode is None:
        return None
    return {"kind": "grid", "n_blocks": int(C.shape[0]), "centers": C,
            "csq": np.ascontiguousarray(np.sum(C ** 2, axis=1)), "dot_mode": mode}


def _describe_tree(tree, X):
    if X.shape[1] < 2:
        return None
    n_nodes = len(tree.child)
    center = np.ascontiguousarray(np.array(tree.center, dtype=np.float64).reshape(n_nodes, 2))
    direction = np.ascontiguousarray(np.array(tree.direction, dtype=np.float64).reshape(n_nodes, 2))
    cut = np.ascontiguousarray(np.array(tree.cut, dtype=np.float64).reshape(n_nodes))
    child = np.ascontiguousarray(np.array(tree.child, dtype=np.int32).reshape(n_nodes, 2))
    mode = 0
    if n_nodes > 0:
        P = np.array(X[:, :2], dtype=np.float64, copy=True)
        P[:, 0] = (X[:, 0] + 22) % 360 - 22
        modes = set()
        for node in sorted(set([tree._root, n_nodes - 1])):
            if node < 0:
                continue
            for rows in (np.arange(P.shape[0]), _sample(P.shape[0], 37)):
                V = P[rows] - center[node]
                got = np.dot(V, direction[node])            # gemv, as in PDTree.recluster
                si = _sample(len(rows), N_PROBE)
                m = _pick_mode(V[si], direction[node][None, :], got[si][:, None])
                modes.add(m)
        if len(modes) != 1 or None in modes:
            return None
        mode = modes.pop()
    return {"kind": "tree", "n_nodes": n_nodes, "center": center, "direction": direction, "cut": cut,
            "child": child, "root": int(tree._root), "n_leaves": len(tree.leaves),
            "wrap_add": 22.0, "wrap_mod": 360.0, "dot_mode": mode}


def describe(block_fn, X):
    """Partitioner description for the C-ABI, or None when ``block_fn`` must run on the host."""
    owner = getattr(block_fn, "__self__", None)
    if isinstance(owner, Blocker) and getattr(block_fn, "__name__", "") == "block_clusters":
        return _describe_grid(owner, X)
    tree = getattr(block_fn, "tree", None)
    if isinstance(tree, PDTree):
        return _describe_tree(tree, X)
    return None
