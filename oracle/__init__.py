"""CPU oracle for the GPRF llgrad hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gprf_b200/`` may import this
package: it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` have something to
check the CUDA path against (and to time on the host cores).

It is a Python-3 / numpy / scipy restatement of the reference's algorithm
(``/root/reference/gprf.py``, ``gpy_linalg.py``, ``block_clustering.py``,
``pdtree_clustering.py``, ``synthetic.py``, ``gprfopt.py``).  The reference
itself cannot be imported (Python 2 syntax, ``scipy.weave``, un-vendored
``treegp`` @ a0aa7ae65a4b9144a499016bbf0ccaf0c611cc0d - see DESIGN.md).

Pinning status
--------------
* euclidean + SE kernel, objective: PINNED against the golden objective values
  shipped in ``/root/reference/gprf_results.tgz`` (``tests/golden/``).
* gradients: no golden values exist in the reference; pinned by central
  finite differences of the pinned objective.
* lld + Matern-3/2 kernel: PARITY UNPINNED (the seismic data blob and the
  treegp source are absent); only the great-circle distance doctests of
  ``run_seismic.py:24-33`` pin part of it.
"""
