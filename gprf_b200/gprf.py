"""``GPRF`` - the reference's objective class (gprf.py:83-331), CUDA-backed.

Same constructor, mutators, attributes and return conventions as the
reference so that its L-BFGS drivers (gprfopt.py:377-417,
run_seismic.py:121-199) drive it unchanged:

    gprf.update_X(XX)
    ll, gX, gC = gprf.llgrad(local=True, grad_X=True, grad_cov=True, parallel=False)

``ll`` is a float, ``gX`` an (n, dx) array or ``zeros((0, 0))``, ``gC`` a
(1, ncov) array or ``zeros((0, 0))`` (gprf.py:275,288,291).  All numerics run
in libgprf_b200.so (hand-written sm_100a kernels); this file only marshals
numpy arrays across the C-ABI.  There is no CPU fallback.

Block membership (``block_idxs = block_fn(X)``, recomputed by every
``update_X``, gprf.py:169-174) is evaluated on the GPU when ``block_fn`` is one
of this package's partitioners (``Blocker.block_clusters`` or the ``reblock``
closure of ``pdtree_cluster``) AND the device result has been proven identical
to the host numpy result on the construction-time X (see ``partition_probe``);
any other ``block_fn`` is called on the host exactly like the reference does.
"""
import ctypes as C
import os
import warnings
from collections import defaultdict

import numpy as np
from numpy.linalg import LinAlgError

from . import _lib
from . import partition_probe
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


def _edge_array(edges):
    if len(edges) == 0:
        return np.zeros((0, 2), dtype=np.int32)
    return np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))


class _EdgeList(list):
    """The list behind ``GPRF.neighbors``: an ordinary list of (i, j) tuples that counts its
    mutations, so that the per-evaluation check "is the device's edge list still current?" is O(1)
    and an in-place edit (``gprf.neighbors[k] = ...``, ``.append``, ...) still reaches the device.
    (Rebuilding an int32 array from 342 tuples and comparing its bytes cost 135 us per evaluation.)"""

    def __init__(self, *args):
        super(_EdgeList, self).__init__(*args)
        self.version = 0

    def _bump(name):          # noqa: N805
        base = getattr(list, name)

        def method(self, *args, **kwargs):
            self.version += 1
            return base(self, *args, **kwargs)
        method.__name__ = name
        return method

    for _name in ("__setitem__", "__delitem__", "__iadd__", "__imul__", "append", "extend", "insert", "pop",
                  "remove", "clear", "sort", "reverse"):
        locals()[_name] = _bump(_name)
    del _name, _bump

    def __reduce__(self):     # pickles as a plain list
        return (list, (list(self),))


class GPRF(object):

    @property
    def neighbors(self):
        return self._neighbors

    @neighbors.setter
    def neighbors(self, value):
        self._neighbors = value if isinstance(value, _EdgeList) else _EdgeList(value)

    def __init__(self, X, Y, block_fn, cov, noise_var, kernelized=False, dy=None,
                 neighbor_threshold=1e-3, nonstationary=False, nonstationary_prec=False,
                 block_idxs=None, neighbors=None, device=0, unit_shard=None, device_blocks=True):
        """gprf.py:85-117.  Extra keywords: ``device`` (CUDA ordinal), ``unit_shard`` =
        (rank, world) to evaluate only this rank's share of the units (multi-GPU; see
        gprf_b200.dist) and ``device_blocks`` (allow the on-device partitioner)."""
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
        self.device_blocks = device_blocks
        # runtime guard of the on-device re-blocking: every k-th re-blocked evaluation is checked
        # against block_fn(X) on the host (0 = only the construction-time proof)
        self.verify_reblock_every = int(os.environ.get("GPRF_VERIFY_REBLOCK", "0"))
        self._reblock_count = 0
        self._lib = None
        self._h = None
        self._device_part = None
        self._open()
        if block_idxs is None:
            block_idxs = block_fn(X)
        self.block_idxs = block_idxs
        if neighbors is not None:
            self.neighbors = neighbors
        else:
            self.compute_neighbors(threshold=neighbor_threshold)
        self.compute_neighbor_count()
        self.neighbor_dict = symmetrize_neighbors(self.neighbors)
        if device_blocks and block_fn is not None:
            self._setup_device_partitioner()

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
        self._shape = (int(X.shape[0]), int(X.shape[1]))
        self._edges_key = None
        self._blocks_key = None
        self._blocks_stale = False

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

    def _Xc(self, X=None):
        """C-contiguous float64 view of X, checked against the (n, dx) the native handle was created
        with (the library copies exactly n*dx doubles in and out)."""
        Xc = np.ascontiguousarray(self.X if X is None else X, dtype=np.float64)
        if Xc.ndim != 2 or Xc.shape != self._shape:
            raise ValueError("X has shape %s, the GPRF was built for %s" % (Xc.shape, self._shape))
        return Xc

    def _points(self, X):
        """Free-standing point set for the kernel wrappers: (m, dx) with the handle's dx."""
        Xc = np.ascontiguousarray(X, dtype=np.float64)
        if Xc.ndim != 2 or Xc.shape[1] != self._shape[1]:
            raise ValueError("points have shape %s, expected (m, %d)" % (Xc.shape, self._shape[1]))
        return Xc

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
        d["_block_idxs"] = self.block_idxs          # materialise device-held blocks
        for k in ("_lib", "_h", "_edges_key", "_blocks_key", "_Yc", "_keep_edges", "_keep_blocks",
                  "_device_part", "_blocks_stale", "_prior_const"):
            d.pop(k, None)
        return d

    def __setstate__(self, d):
        if "neighbors" in d:                       # pickles written before ``neighbors`` became a property
            d["_neighbors"] = d.pop("neighbors")
        self.__dict__ = d
        if "_neighbors" in d:
            self._neighbors = _EdgeList(d["_neighbors"])
        self._lib = None
        self._h = None
        self._device_part = None
        self.__dict__.setdefault("verify_reblock_every", 0)
        self.__dict__.setdefault("_reblock_count", 0)
        self._open()
        if self.device_blocks and self.block_fn is not None:
            self._setup_device_partitioner()

    # -- block membership ------------------------------------------------------
    @property
    def block_idxs(self):
        """List of index arrays, one per block (gprf.py:98-100).  When membership lives on
        the device it is downloaded on first access."""
        if self._block_idxs is None:
            if self._blocks_stale:
                Xc = self._Xc()
                self._check(self._lib.gprf_reblock(self._h, _lib.ptr(Xc)))
                self._blocks_stale = False
            B = C.c_int()
            plen = C.c_longlong()
            self._check(self._lib.gprf_block_count(self._h, C.byref(B), C.byref(plen)))
            ptr = np.zeros(B.value + 1, dtype=np.int64)
            perm = np.zeros(plen.value, dtype=np.int64)
            self._check(self._lib.gprf_get_blocks(self._h, _lib.ptr(ptr), _lib.ptr(perm)))
            self._block_idxs = [perm[ptr[b]:ptr[b + 1]] for b in range(B.value)]
            self._blocks_key = id(self._block_idxs)      # already on the device
            self._keep_blocks = self._block_idxs
        return self._block_idxs

    @block_idxs.setter
    def block_idxs(self, value):
        self._block_idxs = value
        self._blocks_stale = False
        self.n_blocks = len(value)

    def _setup_device_partitioner(self):
        """Enable on-device block assignment iff it reproduces the host partition of the
        current X bit for bit (and the BLAS rounding mode could be identified)."""
        spec = partition_probe.describe(self.block_fn, np.asarray(self.X, dtype=np.float64))
        if spec is None:
            return
        if spec["kind"] == "grid":
            rc = self._lib.gprf_set_grid_partitioner(self._h, spec["n_blocks"], _lib.ptr(spec["centers"]),
                                                     _lib.ptr(spec["csq"]), spec["dot_mode"])
        else:
            rc = self._lib.gprf_set_tree_partitioner(self._h, spec["n_nodes"], _lib.ptr(spec["center"]),
                                                     _lib.ptr(spec["direction"]), _lib.ptr(spec["cut"]),
                                                     _lib.ptr(spec["child"]), spec["root"], spec["n_leaves"],
                                                     spec["wrap_add"], spec["wrap_mod"], spec["dot_mode"])
        self._check(rc)
        # proof on the actual data: device partition == host partition
        host = self.block_fn(self.X)
        Xc = self._Xc()
        self._check(self._lib.gprf_reblock(self._h, _lib.ptr(Xc)))
        saved = self._block_idxs
        self._block_idxs = None
        dev = self.block_idxs
        same = len(dev) == len(host) and all(np.array_equal(a, b) for a, b in zip(dev, host))
        self._block_idxs = saved
        self._blocks_key = None                   # device now holds `dev`, not `saved`
        if same:
            self._device_part = spec["kind"]

    # -- parameters --------------------------------------------------------
    def _theta(self):
        if len(self.cov.wfn_params) != 1:
            raise ValueError("gradient computation currently assumes just a single scaling parameter for "
                             "weight function, but currently wfn_params=%s" % self.cov.wfn_params)
        return np.ascontiguousarray(np.concatenate([[self.noise_var], np.asarray(self.cov.wfn_params, dtype=float),
                                                    np.asarray(self.cov.dfn_params, dtype=float)]), dtype=np.float64)

    def _device_blocks_match_host(self):
        """Runtime guard of the on-device re-blocking (block_clustering.py:17-26,
        pdtree_clustering.py:65-77): the device-held membership of the current X against block_fn(X)."""
        self._block_idxs = None
        dev = self.block_idxs
        host = self.block_fn(self.X)
        return len(dev) == len(host) and all(np.array_equal(a, b) for a, b in zip(dev, host))

    def update_covs(self, covs):
        """gprf.py:160-167."""
        nv, sv = covs[0, :2]
        self.cov = GPCov(wfn_params=[sv], dfn_params=covs[0, 2:], dfn_str=self.cov.dfn_str,
                         wfn_str=self.cov.wfn_str)
        self.noise_var = nv

    def update_X(self, new_X, update_blocks=True, recompute_neighbors=False):
        """gprf.py:169-174."""
        if np.shape(new_X) != self._shape:
            raise ValueError("new_X has shape %s, the GPRF was built for %s" % (np.shape(new_X), self._shape))
        self.X = new_X
        if self.block_fn is not None:
            if self._device_part is not None:
                self._block_idxs = None           # recomputed on the GPU by the next llgrad
                self._blocks_stale = True
            else:
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

    def _edges_for(self, local):
        if local:
            return self.neighbors
        if getattr(self, "_all_pairs_B", None) != self.n_blocks:
            self._all_pairs = [(i, j) for i in range(self.n_blocks) for j in range(i)]
            self._all_pairs_B = self.n_blocks
        return self._all_pairs

    def _sync_edges(self, edges):
        # O(1) check: ``neighbors`` is an _EdgeList (identity + mutation count); the all-pairs list of
        # local=False is internal and never edited.  The (1 - deg) weights of gprf.py:253-264 are
        # derived on the device from this edge list.
        key = (id(edges), getattr(edges, "version", -1), len(edges))
        if key == self._edges_key:
            return
        e = _edge_array(edges)
        rank, world = self.unit_shard if self.unit_shard is not None else (0, 1)
        self._check(self._lib.gprf_set_edges(self._h, len(e), _lib.ptr(e), int(rank), int(world)))
        self._edges_key = key
        self._keep_edges = (edges, e)

    def _sync_blocks(self):
        """Host-held block lists -> device (no-op when the device already holds them)."""
        if self._blocks_stale or self._block_idxs is None:      # None: membership lives on the device
            return
        key = id(self._block_idxs)
        if key == self._blocks_key:
            return
        ptr, perm = _blocks_to_csr(self._block_idxs)
        self._check(self._lib.gprf_set_blocks(self._h, len(self._block_idxs), _lib.ptr(ptr), _lib.ptr(perm)))
        self._blocks_key = key
        self._keep_blocks = self._block_idxs

    def compute_neighbors(self, threshold=1e-3):
        """gprf.py:119-150: edge (i, j), j < i, iff max |k(X_i, X_j)| / signal_var > threshold."""
        self.neighbors = []
        if threshold == 1.0:
            return
        blocks = self.block_idxs
        self._sync_edges(self.neighbors)
        self._sync_blocks()
        B = len(blocks)
        maxk = np.empty((B, B), dtype=np.float64)
        Xc = self._Xc()
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
        Xc = self._Xc()
        if not self._blocks_stale:
            self._sync_blocks()                     # block count first: edges are validated against it
        self._sync_edges(self._edges_for(local))
        th = self._theta()
        ll = C.c_double()
        failed = C.c_int(-1)
        gX = np.empty(Xc.shape, dtype=np.float64) if grad_X else None
        gC = np.empty(len(th), dtype=np.float64) if grad_cov else None
        reblocked = self._blocks_stale
        if reblocked:
            fn = self._lib.gprf_llgrad_reblock      # update_X's re-blocking happens on the GPU
            self._blocks_stale = False
            self._blocks_key = None
        else:
            fn = self._lib.gprf_llgrad
        rc = fn(self._h, _lib.ptr(Xc), _lib.ptr(th), len(th), int(grad_X), int(grad_cov),
                C.byref(ll), _lib.ptr(gX), _lib.ptr(gC), C.byref(failed))
        self._check(rc, failed.value)
        if reblocked and self.verify_reblock_every > 0:
            self._reblock_count += 1
            if self._reblock_count % self.verify_reblock_every == 0 and not self._device_blocks_match_host():
                # the device partition is no longer the reference's: from here on block_fn runs on the
                # host (block assignment only - the evaluation itself stays on the GPU), and this
                # evaluation is redone with the host's blocks
                warnings.warn("gprf_b200: device re-blocking differs from block_fn(X); block assignment "
                              "falls back to the host block_fn", RuntimeWarning)
                self._device_part = None
                self.block_idxs = self.block_fn(self.X)
                return self.llgrad(parallel=parallel, local=local, **kwargs)
        gradX = gX if grad_X else np.zeros((0, 0))
        gradCov = gC.reshape((1, -1)) if grad_cov else np.zeros((0, 0))
        return np.float64(ll.value), gradX, gradCov

    def llgrad_device(self, X_dev_ptr, out_dev_ptr, stream_ptr=0, local=True, grad_X=False, grad_cov=False,
                      reblock=False, status_dev_ptr=None):
        """Device-resident variant (gprf_llgrad_device): X already in HBM at ``X_dev_ptr``
        (n x dx doubles), results left in HBM at ``out_dev_ptr`` as
        [ll, grad_theta (5, zero padded), gradX (n*dx)].  Pointers are plain integers
        (e.g. ``tensor.data_ptr()``, ``torch.cuda.current_stream().cuda_stream``).
        ``reblock`` recomputes block membership from the device X first (needs the
        on-device partitioner); otherwise the current blocks are used.
        ``status_dev_ptr`` (a device double): no host round trip - on the resident path the launches
        are only enqueued and the evaluation's status (0 = ok) is written there by the device
        (gprf_llgrad_device_nosync); returns True when the evaluation is still in flight.  The caller
        synchronises, reads the status (e.g. after an all-reduce that carried it) and repeats the call
        without ``status_dev_ptr`` when it is non-zero."""
        self._sync_edges(self._edges_for(local))
        if reblock:
            if self._device_part is None:
                raise RuntimeError("reblock=True needs the on-device partitioner")
            self._check(self._lib.gprf_reblock_device(self._h, C.c_void_p(X_dev_ptr), C.c_void_p(stream_ptr)))
            self._block_idxs = None
            self._blocks_stale = False
            self._blocks_key = None
        elif self._blocks_stale:
            _ = self.block_idxs
        else:
            self._sync_blocks()
        th = self._theta()
        failed = C.c_int(-1)
        if status_dev_ptr is not None:
            enq = C.c_int(0)
            rc = self._lib.gprf_llgrad_device_nosync(self._h, C.c_void_p(X_dev_ptr), _lib.ptr(th), len(th), int(grad_X),
                                                     int(grad_cov), C.c_void_p(out_dev_ptr), C.c_void_p(status_dev_ptr),
                                                     C.c_void_p(stream_ptr), C.byref(enq), C.byref(failed))
            self._check(rc, failed.value)
            return bool(enq.value)
        rc = self._lib.gprf_llgrad_device(self._h, C.c_void_p(X_dev_ptr), _lib.ptr(th), len(th), int(grad_X),
                                          int(grad_cov), C.c_void_p(out_dev_ptr), C.c_void_p(stream_ptr),
                                          C.byref(failed))
        self._check(rc, failed.value)
        return False

    def unit_results(self):
        """Per-unit log-likelihoods and applied jitter of the last evaluation
        (units: blocks 0..B-1, then edges in ``neighbors`` order)."""
        U = self.n_blocks + len(self._keep_edges[1])
        lls = np.zeros(U)
        jit = np.zeros(U)
        self._check(self._lib.gprf_unit_results(self._h, _lib.ptr(lls), _lib.ptr(jit)))
        return lls, jit

    def last_timing(self):
        ms = C.c_float()
        n = C.c_int()
        self._lib.gprf_last_timing(self._h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def set_keep_kinv(self, on=True):
        """Also store K^-1 of every unit (lower triangle of its working matrix, see gprf_debug_unit)."""
        self._check(self._lib.gprf_set_keep_kinv(self._h, int(bool(on))))

    def set_fused_nt(self, nt):
        """Units of up to `nt` 64-point tiles use the fused one-CTA-per-unit kernel (0: never)."""
        self._check(self._lib.gprf_set_fused_nt(self._h, int(nt)))

    def set_factor_reuse(self, on=True):
        """Pairs read block i's factor tiles instead of recomputing them (default on; bit-identical).
        ``on=2`` also lets fused pairs reuse when there are only a few of them (include/gprf_b200.h)."""
        self._check(self._lib.gprf_set_factor_reuse(self._h, 2 if on == 2 else int(bool(on))))

    def factor_reuse_stats(self):
        """(pair units reusing their parent's factor, tile tasks saved) for the current structure."""
        nu = C.c_int()
        ntl = C.c_longlong()
        self._check(self._lib.gprf_factor_reuse_stats(self._h, C.byref(nu), C.byref(ntl)))
        return nu.value, ntl.value

    # -- resident (shared-memory) unit path: switches and introspection ----------------------
    def set_resident(self, on=True):
        """Small-block structures are evaluated by the resident kernels (one CTA per unit, the pair
        factorisations reuse block i's factor; include/gprf_b200.h).  ``on=False``: tile pipeline only."""
        self._check(self._lib.gprf_set_resident(self._h, int(bool(on))))

    def resident_stats(self):
        """(evaluations tried on the resident path, of which re-run by the tile pipeline, last status)."""
        ev, fb, st = C.c_longlong(), C.c_longlong(), C.c_int()
        self._check(self._lib.gprf_resident_stats(self._h, C.byref(ev), C.byref(fb), C.byref(st)))
        return ev.value, fb.value, st.value

    @staticmethod
    def _resident_layout():
        lib = _lib.load()
        v = np.zeros(16, dtype=np.int64)
        m = lib.gprf_resident_layout(v.ctypes.data_as(C.POINTER(C.c_longlong)), 16)
        names = ["MAXB", "NYB", "BLK", "EXP_W", "EXP_KINV", "EXP_ZY", "EXP_AROW", "EXP_SCAL", "EXP_STRIDE",
                 "GX_STRIDE", "CAP_DOUBLES", "EXP_KSAVE"]
        return dict(zip(names, (int(x) for x in v[:m])))

    @staticmethod
    def _unswizzle(blk):
        """64 doubles of one packed 8x8 block -> (8, 8) array (resident.cuh: sw_off)."""
        out = np.empty((8, 8))
        for r in range(8):
            for c in range(8):
                out[r, c] = blk[((r ^ ((r >> 1) & 1)) << 3) + (c ^ (r & 4))]
        return out

    def resident_debug(self, unit, phase):
        """Select the (unit, phase) whose shared-memory matrices the next evaluation dumps."""
        self._check(self._lib.gprf_set_resident_debug(self._h, int(unit), int(phase)))

    def resident_dump(self):
        """(R1, R2) as dense 160 x 160 arrays, from the last evaluation with resident_debug set."""
        out = np.zeros((2, 160, 160))
        self._check(self._lib.gprf_get_resident_debug(self._h, _lib.ptr(out), -1, None, -1, None))
        return out[0], out[1]

    def resident_export(self, block, nb):
        """Export record of a block unit with ``nb`` points: dict(W, Kinv, Z, alpha, logdet, q)."""
        lay = self._resident_layout()
        raw = np.zeros(lay["EXP_STRIDE"])
        self._check(self._lib.gprf_get_resident_debug(self._h, None, int(block), _lib.ptr(raw), -1, None))
        bb = (nb + 7) // 8
        BL, MAXB, NYB = lay["BLK"], lay["MAXB"], lay["NYB"]
        nyb = (self._Yc.shape[1] + 7) // 8

        def tri_mat(off):
            M = np.zeros((bb * 8, bb * 8))
            for i in range(bb):
                for j in range(i + 1):
                    o = off + (i * (i + 1) // 2 + j) * BL
                    M[8 * i:8 * i + 8, 8 * j:8 * j + 8] = self._unswizzle(raw[o:o + BL])
            return M
        Z = np.zeros((bb * 8, NYB * 8))
        A = np.zeros((bb * 8, NYB * 8))
        for k in range(bb):
            for y in range(nyb):
                o = lay["EXP_ZY"] + (y * bb + k) * BL
                Z[8 * k:8 * k + 8, 8 * y:8 * y + 8] = self._unswizzle(raw[o:o + BL])
                o = lay["EXP_AROW"] + (k * nyb + y) * BL
                A[8 * k:8 * k + 8, 8 * y:8 * y + 8] = self._unswizzle(raw[o:o + BL])
        return dict(W=tri_mat(lay["EXP_W"]), Kinv=tri_mat(lay["EXP_KINV"]), Z=Z, alpha=A,
                    logdet=raw[lay["EXP_SCAL"]], q=raw[lay["EXP_SCAL"] + 1])

    def resident_unit(self, unit):
        """(ll, grad theta (5), gradX rows in the unit's padded local order (GX_STRIDE/3, 3))."""
        lay = self._resident_layout()
        raw = np.zeros(1 + _lib.MAX_NCOV + lay["GX_STRIDE"])
        self._check(self._lib.gprf_get_resident_debug(self._h, None, -1, None, int(unit), _lib.ptr(raw)))
        return raw[0], raw[1:1 + _lib.MAX_NCOV], raw[1 + _lib.MAX_NCOV:].reshape(-1, 3)

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

    def set_unit_mask(self, mask=None, raw_weights=False):
        """Restrict the evaluations to the units with mask != 0 (units: blocks 0..B-1, then the edges
        in ``neighbors`` order); ``None`` restores all units.  ``raw_weights``: every active unit counts
        once instead of with (1 - deg_i)  (gprf_set_unit_mask)."""
        self._sync_blocks()
        self._sync_edges(self._edges_for(True))
        if mask is None:
            self._check(self._lib.gprf_set_unit_mask(self._h, None, 0, 0))
            return
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        if m.shape != (self.n_blocks + len(self._keep_edges[1]),):
            raise ValueError("mask must have one entry per unit (%d blocks + %d edges)"
                             % (self.n_blocks, len(self._keep_edges[1])))
        self._check(self._lib.gprf_set_unit_mask(self._h, _lib.ptr(m), len(m), int(bool(raw_weights))))

    # -- optimiser glue (gprfopt.py:396-409, run_seismic.py:157-179) --------------------------------------
    def set_x_prior(self, mean, std, grad_scale=None):
        """Independent Gaussian prior on the locations, N(mean_pd, std_d^2) (gprfopt.py:172-182 with
        one std; run_seismic.py:363-371 with one per column).  ``grad_scale``: per-column factor of the
        returned gradient (the seismic driver optimises depth / 100).  ``mean=None`` removes it."""
        if mean is None:
            self._check(self._lib.gprf_set_x_prior(self._h, None, None, None))
            self._prior_const = None
            return
        m = np.ascontiguousarray(mean, dtype=np.float64)
        if m.shape != self._shape:
            raise ValueError("prior mean has shape %s, expected %s" % (m.shape, self._shape))
        dx = self._shape[1]
        sd = np.ascontiguousarray(np.broadcast_to(np.asarray(std, dtype=np.float64), (dx,)))
        gs = np.ascontiguousarray(np.broadcast_to(np.asarray(1.0 if grad_scale is None else grad_scale,
                                                             dtype=np.float64), (dx,)))
        iv = np.ascontiguousarray(1.0 / sd ** 2)
        self._check(self._lib.gprf_set_x_prior(self._h, _lib.ptr(m), _lib.ptr(iv), _lib.ptr(gs)))
        n = self._shape[0]
        self._prior_const = -.5 * n * (dx * np.log(2 * np.pi) + np.sum(np.log(sd ** 2)))

    def neg_objective(self, X, grad_cov=False):
        """One L-BFGS callback evaluation: ``update_X(X)`` + ``llgrad(grad_X=True)`` + the location prior,
        returned as the minimiser wants it, f = -(ll + x_prior(X)) and g = -(gradX + prior gradient) *
        grad_scale, both formed on the device (gprf_neg_objective).  Also returns d ll / d theta
        ((1, ncov) or zeros((0, 0))) for the host's log-theta chain rule."""
        if getattr(self, "_prior_const", None) is None:
            raise RuntimeError("neg_objective needs set_x_prior first")
        self.update_X(X)
        Xc = self._Xc()
        if not self._blocks_stale:
            self._sync_blocks()
        self._sync_edges(self._edges_for(True))
        th = self._theta()
        f = C.c_double()
        failed = C.c_int(-1)
        g = np.empty(Xc.shape, dtype=np.float64)
        gC = np.empty(len(th), dtype=np.float64) if grad_cov else None
        reblock = self._blocks_stale
        if reblock:
            self._blocks_stale = False
            self._blocks_key = None
        rc = self._lib.gprf_neg_objective(self._h, _lib.ptr(Xc), _lib.ptr(th), len(th), int(grad_cov), int(reblock),
                                          C.byref(f), _lib.ptr(g), _lib.ptr(gC), C.byref(failed))
        self._check(rc, failed.value)
        return f.value - self._prior_const, g, (gC.reshape((1, -1)) if grad_cov else np.zeros((0, 0)))

    def _masked_unit(self, unit, rows, **kwargs):
        """One unit of the live structure through the normal evaluation path (unit mask, weight 1)."""
        grad_X, grad_cov = bool(kwargs.get("grad_X", False)), bool(kwargs.get("grad_cov", False))
        mask = np.zeros(self.n_blocks + len(self.neighbors), dtype=np.uint8)
        mask[unit] = 1
        self.set_unit_mask(mask, raw_weights=True)
        try:
            ll, gX, gC = self.llgrad(grad_X=grad_X, grad_cov=grad_cov)
        finally:
            self.set_unit_mask(None)
        return float(ll), (gX[rows] if grad_X else np.zeros(())), (gC.reshape(-1) if grad_cov else np.zeros(()))

    def llgrad_unary(self, i, sparse=False, **kwargs):
        """gprf.py:299-308, on the live structure: the unit mask selects block i, nothing is re-created."""
        idx = self.block_idxs[i]
        if len(idx) == 0:
            return self.gaussian_llgrad(self.X[idx], self.Y[idx], **kwargs)
        return self._masked_unit(int(i), idx, **kwargs)

    def llgrad_joint(self, i, j, sparse=False, **kwargs):
        """gprf.py:310-330.  An edge of ``neighbors`` runs as that pair unit of the live structure
        (block i's factor is shared as in a full evaluation); any other pair goes through
        gaussian_llgrad on the stacked rows."""
        ii, jj = self.block_idxs[i], self.block_idxs[j]
        try:
            e = self.neighbors.index((i, j))
        except ValueError:
            e = -1
        if e < 0 or len(ii) + len(jj) == 0:
            return self.gaussian_llgrad(np.vstack([self.X[ii], self.X[jj]]),
                                        np.vstack([self.Y[ii], self.Y[jj]]), **kwargs)
        return self._masked_unit(self.n_blocks + e, np.concatenate([ii, jj]), **kwargs)

    # -- prediction (gprf.py:593-672) ---------------------------------------------
    def block_precisions(self, Y=None):
        """Per block: (K_b + nv I)^-1 and Alpha_b = K_b^-1 Y_b, computed by the device pipeline
        (jitchol -> triangular inverse -> U U^T, the dpotri route of gpy_linalg.py:219-240).  With the
        GPRF's own Y this runs on the live handle: one evaluation of the block units only (unit mask,
        K^-1 tiles kept), no second context; another Y needs its own upload and therefore its own handle."""
        blocks = self.block_idxs
        if Y is not None and Y is not self.Y:
            sub = GPRF(np.ascontiguousarray(self.X, dtype=np.float64), np.ascontiguousarray(Y, dtype=np.float64), None,
                       self.cov, self.noise_var, block_idxs=blocks, neighbors=[], device=self.device)
            try:
                return sub.block_precisions()
            finally:
                sub.close()
        B = self.n_blocks
        mask = np.zeros(B + len(self.neighbors), dtype=np.uint8)
        mask[:B] = 1
        self.set_unit_mask(mask, raw_weights=True)
        self.set_keep_kinv(True)
        try:
            self.llgrad(grad_X=True)
            Kinvs, Alphas = [], []
            dy = self._Yc.shape[1]
            for b, idx in enumerate(blocks):
                s = len(idx)
                if s == 0:
                    Kinvs.append(np.zeros((0, 0)))
                    Alphas.append(np.zeros((0, dy)))
                    continue
                sz, sp, yr = C.c_int(), C.c_int(), C.c_int()
                self._check(self._lib.gprf_debug_unit(self._h, b, C.byref(sz), C.byref(sp), C.byref(yr), None, None, None))
                M = np.empty((sp.value + yr.value, sp.value))
                Al = np.empty((sp.value, yr.value))
                self._check(self._lib.gprf_debug_unit(self._h, b, None, None, None, _lib.ptr(M), _lib.ptr(Al), None))
                L = np.tril(M[:s, :s])
                Kinvs.append(L + np.tril(L, -1).T)
                Alphas.append(np.array(Al[:s, :dy]))
        finally:
            self.set_keep_kinv(False)
            self.set_unit_mask(None)
        return Kinvs, Alphas

    def train_predictor(self, test_cov=None, Y=None):
        """gprf.py:593-672: returns ``predict(Xstar, test_noise_var=0.0, local=False) -> (mean, cov)``,
        the Bayesian-committee fusion of the GP predictions of the test points' block and its
        neighbour blocks.  The O(b^3) part (per-block inverses and Alpha) runs on the GPU; the
        fusion itself is a handful of ntest x ntest host operations per source block, with the
        cross-covariances evaluated by the device kernel-matrix entry (gprf_kernel_matrix)."""
        Yc = self.Y if Y is None else Y
        block_Kinvs, block_Alphas = self.block_precisions(Y)
        if test_cov is None:
            kern_gp, own = self, False
        else:                                   # the reference builds a dummy VectorTree for test_cov (:602-605)
            dummy = np.zeros((1, np.asarray(self.X).shape[1]))
            kern_gp = GPRF(dummy, np.zeros((1, 1)), None, test_cov, 0.0, block_idxs=[np.arange(1)], neighbors=[],
                           device=self.device)
            own = True
        gp = self
        dy = np.asarray(Yc).shape[1]

        def kernel_fn(A, B):                    # Kstar, Kss: the TRAINING covariance (predict_tree, gprf.py:649-654)
            return gp.kernel(A, B)              # X2 given: cross kernel without noise (gprf.py:341-342)

        def prior_kernel_fn(A, B):              # test_cov enters the prior covariance only (gprf.py:599-605,621)
            return kern_gp.kernel(A, B)

        def predict(Xstar, test_noise_var=0.0, local=False):
            Xstar = np.ascontiguousarray(Xstar, dtype=np.float64)
            prior_cov = prior_kernel_fn(Xstar, Xstar) + np.eye(Xstar.shape[0]) * test_noise_var
            prior_prec = np.linalg.inv(prior_cov)
            prior_mean = np.zeros((Xstar.shape[0], dy))
            source_blocks = set()
            for i, idxs in enumerate(gp.block_fn(Xstar)):
                if len(idxs) == 0:
                    continue
                source_blocks.add(i)
                for j in gp.neighbor_dict[i]:
                    source_blocks.add(j)
            blocks = gp.block_idxs
            Xall = np.asarray(gp.X)
            for i in sorted(source_blocks):
                Kstar = kernel_fn(Xstar, Xall[blocks[i]])
                Kss = kernel_fn(Xstar, Xstar)
                if test_noise_var > 0:
                    Kss = Kss + np.eye(Kss.shape[0]) * gp.noise_var       # gprf.py:653-655
                mean = np.dot(Kstar, block_Alphas[i])
                cov = Kss - np.dot(Kstar, np.dot(block_Kinvs[i], Kstar.T))
                prec = np.linalg.inv(cov)
                prior_mean += np.dot(prec, mean)
                prior_prec += prec - np.linalg.inv(Kss)
            final_cov = np.linalg.inv(prior_prec)
            return np.dot(final_cov, prior_mean), final_cov

        predict._kernel_owner = kern_gp if own else None     # keeps the helper handle alive
        return predict

    # -- kernel wrappers (gprf.py:333-375) --------------------------------------
    def kernel(self, X, X2=None):
        X1 = self._points(X)
        Xb = None if X2 is None else self._points(X2)
        n2 = X1.shape[0] if Xb is None else Xb.shape[0]
        K = np.empty((X1.shape[0], n2), dtype=np.float64)
        th = self._theta()
        self._check(self._lib.gprf_kernel_matrix(self._h, _lib.ptr(X1), X1.shape[0], _lib.ptr(Xb), n2,
                                                 _lib.ptr(th), len(th), _lib.ptr(K)))
        return K

    def dKdx(self, X, p, i, return_vec=False, dKv=None):
        """gprf.py:345-360: derivative of kernel(X, X) wrt coordinate ``i`` of point ``p``.
        ``return_vec``: row p of d k(x_p, .)/d x_{p,i} (entry p zero), written into ``dKv`` when
        given; otherwise the symmetric matrix whose row and column p hold that vector."""
        Xc = self._points(X)
        n = Xc.shape[0]
        th = self._theta()
        row = np.empty(n, dtype=np.float64)
        self._check(self._lib.gprf_kernel_deriv(self._h, _lib.ptr(Xc), n, _lib.ptr(th), len(th), 0, int(p), int(i),
                                                _lib.ptr(row)))
        if return_vec:
            if dKv is None:
                return row
            dKv[:] = row
            return dKv
        dK = np.zeros((n, n))
        dK[p, :] = row
        return dK + dK.T

    def dKdi(self, X1, i):
        """gprf.py:362-375: d(kernel(X1) + nv I)/d theta_i for theta = [nv, s2, lengthscales...]."""
        Xc = self._points(X1)
        n = Xc.shape[0]
        if i == 0:
            return np.eye(n)
        if i == 1:
            if len(self.cov.wfn_params) != 1:
                raise ValueError("gradient computation currently assumes just a single scaling parameter for "
                                 "weight function, but currently wfn_params=%s" % self.cov.wfn_params)
            return self.kernel(Xc, Xc) / self.cov.wfn_params[0]
        th = self._theta()
        out = np.empty((n, n), dtype=np.float64)
        self._check(self._lib.gprf_kernel_deriv(self._h, _lib.ptr(Xc), n, _lib.ptr(th), len(th), 1, 0, int(i) - 2,
                                                _lib.ptr(out)))
        return out

    def subset_llgrad(self, blocks):
        """gprf.py:182-204: objective of the sub-field induced by ``blocks`` (unaries of the subset,
        pairs inside it, neighbour counts restricted to it).  One masked device evaluation."""
        blocks = [int(b) for b in blocks]
        bset = set(blocks)
        inside = [(i, j) for (i, j) in self.neighbors if i in bset and j in bset]
        cnt = defaultdict(int)
        for (i, j) in inside:
            cnt[i] += 1
            cnt[j] += 1
        B = self.n_blocks
        mask = np.zeros(B + len(self.neighbors), dtype=np.uint8)
        mask[blocks] = 1
        eidx = [e for e, (i, j) in enumerate(self.neighbors) if i in bset and j in bset]
        mask[[B + e for e in eidx]] = 1
        self.set_unit_mask(mask, raw_weights=True)
        try:
            self.llgrad()
            lls_all, _ = self.unit_results()
        finally:
            self.set_unit_mask(None)
        lls = np.concatenate([lls_all[:B], lls_all[[B + e for e in eidx]]]) if eidx else lls_all[:B]
        B = self.n_blocks
        ll = float(np.sum(lls[B:B + len(inside)])) if inside else 0.0
        ll += float(np.sum([(1 - cnt[b]) * lls[b] for b in blocks]))
        return ll
