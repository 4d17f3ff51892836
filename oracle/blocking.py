"""Point -> block partitioners of the reference (oracle side).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Follows block_clustering.py:4-45 (``pair_distances``, ``Blocker``),
block_clustering.py:48-103 (``cluster_rpc``), pdtree_clustering.py:4-94
(``PDTree``, ``pdtree_cluster``) and gprfopt.py:519-523 (``grid_centers``).

Deliberate deviation (documented in SURVEY.md F8 / DESIGN.md): the grid edge
builder forces the diagonal of the centre-distance matrix to zero before
looking for the smallest positive distances.  The reference relies on
``sqrt(a - 2ab + b)`` rounding to exactly 0 on the diagonal, which is
BLAS-dependent; the golden objective values need the intended 8-connected
edge set (342 edges on the 10x10 grid).
"""
import numpy as np


def pair_distances(A, B):
    """Expanded-form Euclidean distances, same op order as block_clustering.py:4-5."""
    a2 = np.sum(A ** 2, axis=1)
    b2 = np.sum(B ** 2, axis=1)
    with np.errstate(invalid="ignore"):      # tiny negative roundoff -> NaN, as in the reference
        return np.sqrt(np.outer(a2, np.ones(B.shape[0])) - 2 * np.dot(A, B.T)
                       + np.outer(np.ones(A.shape[0]), b2))


def grid_centers(nblocks):
    """ceil(sqrt(nblocks))^2 centres at odd multiples of 1/(2m) (gprfopt.py:519-523)."""
    m = int(np.ceil(np.sqrt(nblocks)))
    pts = np.linspace(0, 1, 2 * m + 1)[1::2]
    return [np.array((xx, yy)) for xx in pts for yy in pts]


class Blocker(object):
    def __init__(self, block_centers):
        self.block_centers = np.asarray(block_centers)
        self.n_blocks = len(block_centers)

    def block_clusters(self, X):
        """Nearest-centre assignment; ascending indices per block (:17-26)."""
        owner = np.argmin(pair_distances(X, self.block_centers), axis=1)
        everyone = np.arange(len(X))
        return [everyone[owner == b] for b in range(self.n_blocks)]

    def neighbors(self, diag_connections=True):
        """Edges (i, j), j < i, between grid-adjacent centres (:28-45)."""
        if self.n_blocks <= 1:
            return []
        D = pair_distances(self.block_centers, self.block_centers)
        np.fill_diagonal(D, 0.0)                      # F8 fix
        pos = D[D > 0]
        near = np.min(pos) + 1e-6
        diag = np.min(pos[pos > near]) + 1e-6
        cut = diag if diag_connections else near
        return [(i, j) for i in range(self.n_blocks) for j in range(i) if D[i, j] < cut]


def cluster_rpc(X, idxs, target_size, fixed_split=None):
    """Random-projection median splits (block_clustering.py:48-103).

    Consumes ``np.random.choice`` exactly as the reference does when
    ``fixed_split`` is None.
    """
    n = len(idxs)
    if fixed_split is not None and len(fixed_split) == 0:
        return [idxs], ()
    if fixed_split is None:
        if n < target_size:
            return [idxs], ()
        i1 = np.random.choice(idxs)
        i2 = i1
        while i2 == i1:
            i2 = np.random.choice(idxs)
        origin = X[i2, :]
        axis = X[i1, :] - origin
        axis = axis / np.linalg.norm(axis)
        sub1 = sub2 = None
    else:
        (axis, origin), sub1, sub2 = fixed_split
    if n > 0:
        proj = np.array([np.dot(X[i, :] - origin, axis) for i in idxs])
        med = np.median(proj)
        left, right = idxs[proj < med], idxs[proj >= med]
    else:
        left, right = idxs[:0], idxs[:0]
    L1, s1 = cluster_rpc(X, left, target_size, sub1)
    L2, s2 = cluster_rpc(X, right, target_size, sub2)
    return L1 + L2, ((axis, origin), s1, s2)


class PDTree(object):
    """Principal-direction median-split tree (pdtree_clustering.py:4-77)."""

    def __init__(self, X, minsize):
        self.X = X
        self.root = self._grow(np.arange(len(X)), minsize)

    def _grow(self, idx, minsize):
        if len(idx) < minsize:
            return ("leaf", idx)
        pts = self.X[idx]
        mu = np.mean(pts, axis=0)
        pts = pts - mu
        ev, evec = np.linalg.eig(np.dot(pts.T, pts))
        direction = evec[:, np.argmax(ev)]
        proj = np.dot(pts, direction)
        cut = np.median(proj)
        return ("node", direction, mu, cut,
                self._grow(idx[proj < cut], minsize),
                self._grow(idx[proj >= cut], minsize))

    def leaf_idx(self):
        out = []

        def walk(nd):
            if nd[0] == "leaf":
                out.append(nd[1])
            else:
                walk(nd[4])
                walk(nd[5])
        walk(self.root)
        return out

    def recluster(self, X):
        out = []

        def walk(nd, idx):
            if nd[0] == "leaf":
                out.append(idx)
                return
            _, direction, mu, cut, lo, hi = nd
            proj = np.dot(X[idx] - mu, direction)
            walk(lo, idx[proj < cut])
            walk(hi, idx[proj >= cut])
        walk(self.root, np.arange(len(X)))
        return out


def pdtree_cluster(X, blocksize=300):
    """Tree on (wrapped lon, lat); returns (leaf index lists, reblock fn) (:79-94)."""
    P = X[:, :2].copy()
    P[:, 0] = (P[:, 0] + 22) % 360 - 22
    tree = PDTree(P, minsize=blocksize)

    def reblock(XX):
        Q = XX[:, :2].copy()
        Q[:, 0] = (XX[:, 0] + 22) % 360 - 22
        return tree.recluster(Q)

    return tree.leaf_idx(), reblock
