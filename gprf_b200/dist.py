"""Multi-GPU evaluation: one process per GPU, units sharded across ranks.

The reference's only parallelism is data-parallel over independent units
(``Pool.map_async`` over blocks, then over edges - gprf.py:218-233) followed by
a serial reduction in the parent (gprf.py:245-291).  Here every rank holds X, Y
and the full structure (X is n*dx*8 bytes - 3.2 MB at n = 200k - so replicating
it costs less than any halo bookkeeping), evaluates its share of the units on
its own GPU and the partial ``[ll, grad_theta, gradX]`` vectors are summed with
ONE all-reduce over NCCL/NVLink.  No other data-path collective exists.

``shard_units`` is pure host logic (tested with gloo on CPU).
"""
import ctypes as C
import heapq

import numpy as np

from . import _lib
from .gprf import GPRF, LinAlgError

COST_DY = 50.0


def unit_costs(block_ptr, edges, nominal=False):
    """Work model W(s) = s^3 + 4 s^2 dy per unit.  ``nominal``: every block
    counts as 100 points - the split the library uses for small-block structures, whose block
    sizes never reach the host (resident path; ``shard_sizes`` in gprf_lib.cu)."""
    sizes = np.diff(np.asarray(block_ptr, dtype=np.int64)).astype(np.float64)
    if nominal:
        sizes = np.full_like(sizes, 100.0)
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    s = np.concatenate([sizes, sizes[e[:, 0]] + sizes[e[:, 1]]]) if len(e) else sizes
    return s ** 3 + 4.0 * COST_DY * s ** 2


def shard_owners(block_ptr, edges, world, nominal=False):
    """Owner rank of every unit (blocks, then edges).

    The unit of assignment is a *group*: block i together with every edge (i, j) whose stacked
    rows start with block i (gprf.py:310-330), so a pair always finds its parent block's factor
    on its own rank (factor reuse, ``gprf_set_factor_reuse``).  Groups are placed by a
    longest-processing-time greedy on the work model; deterministic, so every rank derives
    the same global assignment without communication (same rule as ``lpt_mask`` in C++).
    """
    cost = unit_costs(block_ptr, edges, nominal)
    B = len(block_ptr) - 1
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    group = cost[:B].copy()
    for k in range(len(e)):                       # sequential adds: same rounding as the C++ loop
        group[e[k, 0]] += cost[B + k]
    owner_b = np.zeros(B, dtype=np.int64)
    if world > 1:
        order = np.argsort(-group, kind="stable")
        heap = [(0.0, r) for r in range(world)]
        heapq.heapify(heap)
        for b in order:
            load, r = heapq.heappop(heap)
            owner_b[b] = r
            heapq.heappush(heap, (load + group[b], r))
    return np.concatenate([owner_b, owner_b[e[:, 0]]]) if len(e) else owner_b


def shard_units(block_ptr, edges, rank, world, nominal=False):
    """uint8 mask over units (blocks, then edges): 1 = evaluated on ``rank``."""
    return np.ascontiguousarray((shard_owners(block_ptr, edges, world, nominal) == rank).astype(np.uint8))


def pack(ll, gX, gC, n, dx):
    """[ll, gC(5, zero padded), gX.ravel()] - the layout of gprf_llgrad_device's out buffer."""
    buf = np.zeros(1 + _lib.MAX_NCOV + n * dx, dtype=np.float64)
    buf[0] = ll
    if gC is not None and gC.size:
        buf[1:1 + gC.size] = gC.ravel()
    if gX is not None and gX.size:
        buf[1 + _lib.MAX_NCOV:] = gX.ravel()
    return buf


def unpack(buf, n, dx, ncov, grad_X, grad_cov):
    ll = np.float64(buf[0])
    gC = np.array(buf[1:1 + ncov], dtype=np.float64).reshape(1, -1) if grad_cov else np.zeros((0, 0))
    gX = np.array(buf[1 + _lib.MAX_NCOV:1 + _lib.MAX_NCOV + n * dx], dtype=np.float64).reshape(n, dx) \
        if grad_X else np.zeros((0, 0))
    return ll, gX, gC


class ShardedGPRF(GPRF):
    """GPRF whose ``llgrad`` evaluates this rank's units and all-reduces.

    Requires an initialised ``torch.distributed`` process group (NCCL on GPUs)
    with one process per GPU; the result is identical on every rank.
    """

    def __init__(self, *args, **kwargs):
        import torch
        import torch.distributed as dist
        self._torch = torch
        self._dist = dist
        rank, world = dist.get_rank(), dist.get_world_size()
        kwargs.setdefault("device", torch.cuda.current_device())
        kwargs["unit_shard"] = (rank, world)
        super(ShardedGPRF, self).__init__(*args, **kwargs)
        self._out = None
        self._Xd = None

    def llgrad(self, parallel=False, local=True, **kwargs):
        torch, dist = self._torch, self._dist
        grad_X = bool(kwargs.get("grad_X", False))
        grad_cov = bool(kwargs.get("grad_cov", False))
        n, dx = self.X.shape
        dev = torch.device("cuda", self.device)
        if self._out is None:
            # slot 0 carries the per-rank status so that a failed factorisation on one
            # rank cannot leave the others waiting in the collective
            self._out = torch.empty(2 + _lib.MAX_NCOV + n * dx, dtype=torch.float64, device=dev)
            self._Xd = torch.empty((n, dx), dtype=torch.float64, device=dev)
            self._Xh = torch.empty((n, dx), dtype=torch.float64).pin_memory()
            self._outh = torch.empty(2 + _lib.MAX_NCOV + n * dx, dtype=torch.float64).pin_memory()
        self._Xh.numpy()[...] = self.X
        self._Xd.copy_(self._Xh, non_blocking=True)
        stream = torch.cuda.current_stream(dev)
        used = 2 + _lib.MAX_NCOV + (n * dx if grad_X else 0)
        # Slot 0 carries the status: on the resident path the device writes it (no host round trip before
        # the collective), otherwise the host does.  A non-zero sum after the all-reduce means that some
        # rank needs the jitter rule / the tile pipeline or failed: then EVERY rank repeats the evaluation
        # synchronously (same decision everywhere - the reduced status is the same on every rank).
        reblock = self._blocks_stale and self._device_part is not None

        def evaluate(sync):
            error = None
            try:
                if sync:
                    self.llgrad_device(self._Xd.data_ptr(), self._out.data_ptr() + 8, stream.cuda_stream, local=local,
                                       grad_X=grad_X, grad_cov=grad_cov, reblock=reblock and not self._reblocked)
                    self._out[0] = 0.0
                else:
                    self.llgrad_device(self._Xd.data_ptr(), self._out.data_ptr() + 8, stream.cuda_stream, local=local,
                                       grad_X=grad_X, grad_cov=grad_cov, reblock=reblock,
                                       status_dev_ptr=self._out.data_ptr())
                    self._reblocked = reblock
            except Exception as exc:          # any failure: flag it, still take part in the collective
                error = exc
                self._out[:used].zero_()
                self._out[0] = 1e6
            dist.all_reduce(self._out[:used])
            self._outh[:used].copy_(self._out[:used], non_blocking=True)
            stream.synchronize()
            return error

        # Every rank takes the same path through the collectives: the decisions below depend on the
        # REDUCED status only (>= 1e6: some rank raised; otherwise non-zero: some rank's device status).
        self._reblocked = False
        error = evaluate(sync=False)
        status = self._outh[0].item()
        if 0.0 < status < 1e6:
            error = evaluate(sync=True)
            status = self._outh[0].item()
        if error is not None:
            raise error
        if status != 0.0:
            raise LinAlgError("the evaluation failed on another rank (e.g. a unit was not positive definite)")
        return unpack(self._outh.numpy()[1:], n, dx, 2 + len(self.cov.dfn_params), grad_X, grad_cov)
