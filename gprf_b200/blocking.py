"""Point -> block partitioners (host side, numpy).

Same interface and results as the reference's block_clustering.py /
pdtree_clustering.py / gprfopt.grid_centers; block membership must be
bit-exact with the reference, so the distance expression that decides ties is
kept operation-for-operation (block_clustering.py:4-5) while the bucketing is
done with one stable sort instead of one mask per block.

Deliberate deviation: ``Blocker.neighbors`` zeroes the diagonal of the
centre-distance matrix before searching for the smallest positive distances
 - the reference's result depends on BLAS rounding there, and
its shipped golden objective values need the intended 8-connected grid.
"""
from collections import defaultdict

import numpy as np


def pair_distances(Xi, Xj):
    """block_clustering.py:4-5 - expanded-form distances, identical op order."""
    sq_i = np.sum(Xi ** 2, axis=1)
    sq_j = np.sum(Xj ** 2, axis=1)
    with np.errstate(invalid="ignore"):
        return np.sqrt(np.outer(sq_i, np.ones((Xj.shape[0]),)) - 2 * np.dot(Xi, Xj.T)
                       + np.outer(np.ones((Xi.shape[0]),), sq_j))


def split_by_label(labels, n_groups):
    """Index arrays per label, ascending inside each group (one stable sort)."""
    order = np.argsort(labels, kind="stable")
    bounds = np.searchsorted(labels[order], np.arange(n_groups + 1))
    return [order[bounds[g]:bounds[g + 1]] for g in range(n_groups)]


def grid_centers(nblocks):
    """gprfopt.py:519-523: an m x m grid, m = ceil(sqrt(nblocks)), x-major order."""
    m = int(np.ceil(np.sqrt(nblocks)))
    ticks = np.linspace(0, 1, 2 * m + 1)[1::2]
    return [np.array((gx, gy)) for gx in ticks for gy in ticks]


def symmetrize_neighbors(neighbors):
    """gprf.py:76-81."""
    table = defaultdict(set)
    for (i, j) in neighbors:
        table[i].add(j)
        table[j].add(i)
    return table


class Blocker(object):
    """Nearest-centre blocks (block_clustering.py:7-45)."""

    def __init__(self, block_centers):
        self.block_centers = np.asarray(block_centers)
        self.n_blocks = len(block_centers)

    def get_block(self, X_new):
        return int(np.argmin([np.linalg.norm(X_new - c) for c in self.block_centers]))

    def assign(self, X):
        return np.argmin(pair_distances(X, self.block_centers), axis=1)

    def block_clusters(self, X):
        return split_by_label(self.assign(X), self.n_blocks)

    def neighbors(self, diag_connections=True):
        if self.n_blocks <= 1:
            return []
        cd = pair_distances(self.block_centers, self.block_centers)
        cd[np.diag_indices_from(cd)] = 0.0
        positive = cd[cd > 0]
        axis_dist = positive.min() + 1e-6
        diag_dist = positive[positive > axis_dist].min() + 1e-6
        limit = diag_dist if diag_connections else axis_dist
        ii, jj = np.nonzero(np.tril(cd < limit, -1))
        return [(int(i), int(j)) for i, j in zip(ii, jj)]


def cluster_rpc(X, idxs, target_size, fixed_split=None):
    """Random-projection tree clustering (block_clustering.py:48-103).

    Draws from the global ``np.random`` stream in the reference's order so that
    a seeded run gives the reference's partition.  Returns (clusters, splits);
    ``splits`` replays the same cuts on moved points via ``fixed_split``.
    """
    idxs = np.asarray(idxs)
    if fixed_split is not None and len(fixed_split) == 0:
        return [idxs], ()
    if fixed_split is None:
        if len(idxs) < target_size:
            return [idxs], ()
        a = np.random.choice(idxs)
        b = a
        while b == a:
            b = np.random.choice(idxs)
        base = X[b, :]
        unit = X[a, :] - base
        unit = unit / np.linalg.norm(unit)
        below = above = None
    else:
        (unit, base), below, above = fixed_split
    if len(idxs) > 0:
        coord = np.array([np.dot(X[i, :] - base, unit) for i in idxs])
        cut = np.median(coord)
        lo, hi = idxs[coord < cut], idxs[coord >= cut]
    else:
        lo, hi = idxs, idxs
    c_lo, s_lo = cluster_rpc(X, lo, target_size, below)
    c_hi, s_hi = cluster_rpc(X, hi, target_size, above)
    return c_lo + c_hi, ((unit, base), s_lo, s_hi)


class PDTree(object):
    """Principal-direction divisive partitioning (pdtree_clustering.py:4-77).

    Nodes are kept in flat lists (children indices, cut parameters); leaves are
    numbered left to right, which is the block order the reference returns.
    """

    def __init__(self, X, minsize):
        self.X = X
        self.direction, self.center, self.cut = [], [], []
        self.child = []           # (left, right); negative = -(leaf id) - 1
        self.leaves = []
        self._root = self._build(np.arange(len(X)), minsize)

    def _build(self, idx, minsize):
        if len(idx) < minsize:
            self.leaves.append(idx)
            return -len(self.leaves)
        pts = self.X[idx]
        mu = np.mean(pts, axis=0)
        pts -= mu
        evals, evecs = np.linalg.eig(np.dot(pts.T, pts))
        direction = evecs[:, np.argmax(evals)]
        proj = np.dot(pts, direction)
        cut = np.median(proj)
        me = len(self.child)
        self.direction.append(direction)
        self.center.append(mu)
        self.cut.append(cut)
        self.child.append(None)
        left = self._build(idx[proj < cut], minsize)
        right = self._build(idx[proj >= cut], minsize)
        self.child[me] = (left, right)
        return me

    def leaf_idx(self):
        return list(self.leaves)

    def recluster(self, X):
        out = [None] * len(self.leaves)
        stack = [(self._root, np.arange(len(X)))]
        while stack:
            node, idx = stack.pop()
            if node < 0:
                out[-node - 1] = idx
                continue
            proj = np.dot(X[idx] - self.center[node], self.direction[node])
            left, right = self.child[node]
            stack.append((left, idx[proj < self.cut[node]]))
            stack.append((right, idx[proj >= self.cut[node]]))
        return out


def pdtree_cluster(X, blocksize=300):
    """pdtree_clustering.py:79-94: tree on (lon wrapped at -22, lat)."""
    def wrapped(A):
        P = np.array(A[:, :2], dtype=float, copy=True)
        P[:, 0] = (A[:, 0] + 22) % 360 - 22
        return P

    tree = PDTree(wrapped(X), minsize=blocksize)

    def reblock(XX):
        return tree.recluster(wrapped(XX))

    reblock.tree = tree
    return tree.leaf_idx(), reblock
