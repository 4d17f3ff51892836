"""``GPRF`` - the reference's objective class (gprf.py:83-331), CUDA-backed.

Same constructor, mutators, attributes and return conventions as the
reference so that its L-BFGS drivers (gprfopt.py:377-417,
run_seismic.py:121-199) drive it unchanged:

    ll, gX, gC = gprf.llgrad(local=True, grad_X=True, grad_cov=True, parallel=False)

``ll`` is a float, ``gX`` an (n, dx) array or ``zeros((0, 0))``, ``gC`` a
(1, ncov) array or ``zeros((0, 0))`` (gprf.py:275,288,291).  All numerics run
in libgprf_b200.so (hand-written sm_100a kernels); this file only marshals
numpy arrays across the C-ABI.  There is no CPU fallback.
"""
import ctypes as C
from collections import defaultdict

import numpy as np
from numpy.linalg import LinAlgError

from . import _lib
from .blocking import symmetrize_neighbors
from .cov import GPCov


def _blocks_to_csr(block_idxs):
    sizes = np.fromiter((len(b) for b in block_idxs), dtype=np.int64, count=len(block_idxs))
    ptr = np.zeros(len(block_idxs) + 1, dtype=np.int64)
    np.cumsum(sizes, out=ptr[1:])
    if ptr[-1] > 0:
        perm = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.int64) for b in block_idxs]))
    else:
        perm = np.zeros(0, dtype=np.int64)
    return ptr, perm


class GPRF(object):

    def __init__(self, X, Y, block_fn, cov, noise_var, kernelized=False, dy=None,
                 neighbor_threshold=1e-3, nonstationary=False, nonstationary_prec=False,
                 block_idxs=None, neighbors=None, device=0, unit_shard=None):
        """gprf.py:85-117.  Extra keywords: ``device`` (CUDA ordinal) and
        ``unit_shard`` = (rank, world) to evaluate only this rank's share of the
        units (multi-GPU; see gprf_b200.dist)."""
        if kernelized or nonstationary or nonstationary_prec:
            raise NotImplementedError("kernelized / nonstationary variants are dead code in the reference "
                                      "(gprf.py:104,674-736) and outside the hot path")
        self.X = X
        self.kernelized = False
        self.Y = Y
        self.nonstationary = False
        self.block_fn = block_fn
        self.cov = cov
        self.noise_var = noise_var
        self.neighbor_threshold = neighbor_threshold
        self.device = device
        self.unit_shard = unit_shard
        self._lib = None
        self._h = None
        self._structure_key = None
        self._open()
        if block_idxs is None:
            block_idxs = block_fn(X)
        self.block_idxs = block_idxs
        self.n_blocks = len(block_idxs)
        if neighbors is not None:
            self.neighbors = neighbors
        else:
            self.compute_neighbors(threshold=neighbor_threshold)
        self.compute_neighbor_count()
        self.neighbor_dict = symmetrize_neighbors(self.neighbors)

    # -- native handle ---------------------------------------------------
    def _open(self):
        self._lib = _lib.load()
        dfn_id, wfn_id = self.cov.ids()
        X = np.asarray(self.X)
        self._Yc = np.ascontiguousarray(self.Y, dtype=np.float64)
        h = C.c_void_p()
        rc = self._lib.gprf_create(C.byref(h), int(self.device), X.shape[0], X.shape[1],
                                   self._Yc.shape[1], _lib.ptr(self._Yc), dfn_id, wfn_id)
        self._h = h
        self._check(rc)
        self._structure_key = None

    def _check(self, rc, failed_unit=-1):
        if rc == _lib.OK:
            return
        if rc == _lib.ERR_NOT_PD:
            raise LinAlgError("not positive definite, even with jitter. (unit %d)" % failed_unit)
        if rc == _lib.ERR_NONPOS_DIAG:
            raise LinAlgError("not pd: non-positive diagonal elements")
        detail = self._lib.gprf_last_error(self._h) if self._h else b""
        msg = "%s: %s" % (self._lib.gprf_strerror(rc).decode(), (detail or b"").decode())
        if rc == _lib.ERR_ARG:
            raise ValueError(msg)
        raise RuntimeError(msg)

    def close(self):
        if getattr(self, "_h", None) is not None and self._lib is not None:
            self._lib.gprf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        # gprf.py:738-741 drops the native evaluator when pickling
        d = self.__dict__.copy()
        for k in ("_lib", "_h", "_structure_key", "_Yc"):
            d.pop(k, None)
        return d

    def __setstate__(self, d):
        self.__dict__ = d
        self._lib = None
        self._h = None
        self._open()

    # -- parameters --------------------------------------------------------
    def _theta(self):
        if len(self.cov.wfn_params) != 1:
            raise ValueError("gradient computation currently assumes just a single scaling parameter for "
                             "weight function, but currently wfn_params=%s" % self.cov.wfn_params)
        return np.ascontiguousarray(np.concatenate([[self.noise_var], np.asarray(self.cov.wfn_params, dtype=float),
                                                    np.asarray(self.cov.dfn_params, dtype=float)]), dtype=np.float64)

    def update_covs(self, covs):
        """gprf.py:160-167."""
        nv, sv = covs[0, :2]
        self.cov = GPCov(wfn_params=[sv], dfn_params=covs[0, 2:], dfn_str=self.cov.dfn_str,
                         wfn_str=self.cov.wfn_str)
        self.noise_var = nv

    def update_X(self, new_X, update_blocks=True, recompute_neighbors=False):
        """gprf.py:169-174."""
        self.X = new_X
        if self.block_fn is not None:
            self.block_idxs = self.block_fn(new_X)
        if recompute_neighbors:
            self.compute_neighbors(threshold=self.neighbor_threshold)
            self.compute_neighbor_count()
            self.neighbor_dict = symmetrize_neighbors(self.neighbors)

    def update_X_block(self, i, new_X):
        self.X[self.block_idxs[i]] = new_X

    # -- structure -----------------------------------------------------------
    def compute_neighbor_count(self):
        """gprf.py:152-157."""
        count = defaultdict(int)
        for (i, j) in self.neighbors:
            count[i] += 1
            count[j] += 1
        self.neighbor_count = count

    def _shard_mask(self, n_units, ptr, edges):
        if self.unit_shard is None:
            return None
        from .dist import shard_units
        rank, world = self.unit_shard
        return shard_units(ptr, edges, rank, world)

    def _push_structure(self, edges, force=False):
        """Hand block membership + edge list to the device (gprf_set_structure)."""
        key = (id(self.block_idxs), id(edges), len(edges))
        if not force and key == self._structure_key:
            return
        ptr, perm = _blocks_to_csr(self.block_idxs)
        e = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2)) if len(edges) else \
            np.zeros((0, 2), dtype=np.int32)
        mask = self._shard_mask(len(self.block_idxs) + len(e), ptr, e)
        rc = self._lib.gprf_set_structure(self._h, len(self.block_idxs), _lib.ptr(ptr), _lib.ptr(perm),
                                          len(e), _lib.ptr(e), _lib.ptr(mask))
        self._check(rc)
        self._structure_key = key
        self._keep = (ptr, perm, e, mask, self.block_idxs, edges)   # keep ids alive while cached

    def compute_neighbors(self, threshold=1e-3):
        """gprf.py:119-150: edge (i, j), j < i, iff max |k(X_i, X_j)| / signal_var > threshold."""
        self.neighbors = []
        if threshold == 1.0:
            return
        self._push_structure([], force=True)
        B = len(self.block_idxs)
        maxk = np.empty((B, B), dtype=np.float64)
        Xc = np.ascontiguousarray(self.X, dtype=np.float64)
        th = self._theta()
        self._check(self._lib.gprf_block_max_kernel(self._h, _lib.ptr(Xc), _lib.ptr(th), len(th), _lib.ptr(maxk)))
        self.maxk_cache = maxk
        ii, jj = np.nonzero(np.tril(maxk > threshold, -1))
        self.neighbors = [(int(i), int(j)) for i, j in zip(ii, jj)]

    # -- the hot path ---------------------------------------------------------
    def llgrad(self, parallel=False, local=True, **kwargs):
        """gprf.py:206-296.  ``parallel`` (a fork-per-call Pool in the reference) is
        accepted and ignored: every unit already runs concurrently on the GPU."""
        grad_X = bool(kwargs.get("grad_X", False))
        grad_cov = bool(kwargs.get("grad_cov", False))
        if kwargs.get("sparse", False):
            raise NotImplementedError("sparse (CHOLMOD) unit likelihoods are outside the hot path")
        if local:
            edges = self.neighbors
        else:
            if getattr(self, "_all_pairs_B", None) != self.n_blocks:
                self._all_pairs = [(i, j) for i in range(self.n_blocks) for j in range(i)]
                self._all_pairs_B = self.n_blocks
            edges = self._all_pairs
        self._push_structure(edges)
        Xc = np.ascontiguousarray(self.X, dtype=np.float64)
        th = self._theta()
        ll = C.c_double()
        failed = C.c_int(-1)
        gX = np.empty(Xc.shape, dtype=np.float64) if grad_X else None
        gC = np.empty(len(th), dtype=np.float64) if grad_cov else None
        rc = self._lib.gprf_llgrad(self._h, _lib.ptr(Xc), _lib.ptr(th), len(th), int(grad_X), int(grad_cov),
                                   C.byref(ll), _lib.ptr(gX), _lib.ptr(gC), C.byref(failed))
        self._check(rc, failed.value)
        gradX = gX if grad_X else np.zeros((0, 0))
        gradCov = gC.reshape((1, -1)) if grad_cov else np.zeros((0, 0))
        return np.float64(ll.value), gradX, gradCov

    def llgrad_device(self, X_dev_ptr, out_dev_ptr, stream_ptr=0, local=True, grad_X=False, grad_cov=False):
        """Device-resident variant (gprf_llgrad_device): X already in HBM at ``X_dev_ptr``
        (n x dx doubles), results left in HBM at ``out_dev_ptr`` as
        [ll, grad_theta (5, zero padded), gradX (n*dx)].  Pointers are plain integers
        (e.g. ``tensor.data_ptr()``, ``torch.cuda.current_stream().cuda_stream``).
        Uses the block structure of the last ``update_X`` / constructor."""
        edges = self.neighbors if local else [(i, j) for i in range(self.n_blocks) for j in range(i)]
        self._push_structure(edges)
        th = self._theta()
        failed = C.c_int(-1)
        rc = self._lib.gprf_llgrad_device(self._h, C.c_void_p(X_dev_ptr), _lib.ptr(th), len(th), int(grad_X),
                                          int(grad_cov), C.c_void_p(out_dev_ptr), C.c_void_p(stream_ptr),
                                          C.byref(failed))
        self._check(rc, failed.value)

    def unit_results(self):
        """Per-unit log-likelihoods and applied jitter of the last evaluation
        (units: blocks 0..B-1, then edges in ``neighbors`` order)."""
        U = self.n_blocks + len(self._keep[2])
        lls = np.zeros(U)
        jit = np.zeros(U)
        self._check(self._lib.gprf_unit_results(self._h, _lib.ptr(lls), _lib.ptr(jit)))
        return lls, jit

    def last_timing(self):
        ms = C.c_float()
        n = C.c_int()
        self._lib.gprf_last_timing(self._h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def set_profiling(self, on=True):
        self._lib.gprf_set_profiling(self._h, int(bool(on)))

    def family_timing(self):
        """{family: (ms, launches)} of the last evaluation (needs set_profiling(True))."""
        ms = (C.c_float * _lib.N_FAMILIES)()
        nl = (C.c_int * _lib.N_FAMILIES)()
        self._lib.gprf_family_timing(self._h, ms, nl)
        return dict((self._lib.gprf_family_name(i).decode(), (ms[i], nl[i])) for i in range(_lib.N_FAMILIES))

    # -- single-unit entry points (gprf.py:299-330, 496-591) -------------------
    def gaussian_llgrad(self, X, Y, grad_X=False, grad_cov=False, **_unused):
        n = X.shape[0]
        ncov = 2 + len(self.cov.dfn_params)
        if n == 0:
            return 0.0, (np.zeros(X.shape) if grad_X else np.zeros(())), \
                (np.zeros((ncov,)) if grad_cov else np.zeros(()))
        sub = GPRF(np.ascontiguousarray(X), np.ascontiguousarray(Y), None, self.cov, self.noise_var,
                   block_idxs=[np.arange(n)], neighbors=[], device=self.device)
        try:
            ll, gX, gC = sub.llgrad(grad_X=grad_X, grad_cov=grad_cov)
        finally:
            sub.close()
        return float(ll), (gX if grad_X else np.zeros(())), (gC.reshape(-1) if grad_cov else np.zeros(()))

    def llgrad_unary(self, i, sparse=False, **kwargs):
        idx = self.block_idxs[i]
        return self.gaussian_llgrad(self.X[idx], self.Y[idx], **kwargs)

    def llgrad_joint(self, i, j, sparse=False, **kwargs):
        ii, jj = self.block_idxs[i], self.block_idxs[j]
        return self.gaussian_llgrad(np.vstack([self.X[ii], self.X[jj]]),
                                    np.vstack([self.Y[ii], self.Y[jj]]), **kwargs)

    # -- kernel wrappers (gprf.py:333-375) --------------------------------------
    def kernel(self, X, X2=None):
        X1 = np.ascontiguousarray(X, dtype=np.float64)
        Xb = None if X2 is None else np.ascontiguousarray(X2, dtype=np.float64)
        n2 = X1.shape[0] if Xb is None else Xb.shape[0]
        K = np.empty((X1.shape[0], n2), dtype=np.float64)
        th = self._theta()
        self._check(self._lib.gprf_kernel_matrix(self._h, _lib.ptr(X1), X1.shape[0], _lib.ptr(Xb), n2,
                                                 _lib.ptr(th), len(th), _lib.ptr(K)))
        return K
