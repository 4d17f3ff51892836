"""gprf_b200 - B200-native GPRF objective-and-gradient hot path.

Drop-in for the ``GPRF.llgrad`` path of davmre/gprf (gprf.py:206): the Python
surface of the reference (``GPRF``, ``Blocker``, ``pdtree_cluster``,
``cluster_rpc``, treegp-style ``GPCov``) is kept, the numerics run in
hand-written sm_100a CUDA kernels behind the C-ABI of ``include/gprf_b200.h``
(``libgprf_b200.so``).  There is no CPU fallback: importing ``GPRF`` without
the built library, or evaluating without a CUDA device, raises.
"""
from .cov import GPCov
from .blocking import (Blocker, pair_distances, grid_centers, cluster_rpc, PDTree,
                       pdtree_cluster, symmetrize_neighbors)
from .gprf import GPRF, LinAlgError

__all__ = ["GPRF", "GPCov", "Blocker", "pair_distances", "grid_centers", "cluster_rpc",
           "PDTree", "pdtree_cluster", "symmetrize_neighbors", "LinAlgError"]
