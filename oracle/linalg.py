"""Dense linear algebra of the reference's ``gpy_linalg.py`` (oracle side).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Restates the hot subset only: ``jitchol`` (gpy_linalg.py:77-97), ``pdinv``
(:219-240), ``dpotrs`` (:139-148), ``dpotri``+``symmetrify`` (:150-171,
:410-483).  ``dtrtri`` (:243-253) is computed by the reference's ``pdinv`` but
its result is never used by gprf.py, so it is not restated.
"""
import numpy as np
from scipy.linalg import lapack
from numpy.linalg import LinAlgError


def jitchol(A, maxtries=5, return_jitter=False):
    """Cholesky with the reference's jitter-retry rule (gpy_linalg.py:77-97).

    First attempt: plain dpotrf.  On failure: every diagonal entry must be
    > 0 (else LinAlgError); then up to ``maxtries`` attempts on
    ``A + jitter*I`` with jitter = mean(diag A) * 1e-6 * 10^k, always from
    the *original* A.
    """
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return (L, 0.0) if return_jitter else L
    dg = np.diag(A)
    if np.any(dg <= 0.0):
        raise LinAlgError("not pd: non-positive diagonal elements")
    jitter = dg.mean() * 1e-6
    for _ in range(maxtries):
        if not np.isfinite(jitter):
            break
        L, info = lapack.dpotrf(A + np.eye(A.shape[0]) * jitter, lower=1)
        if info == 0:
            return (L, jitter) if return_jitter else L
        jitter *= 10.0
    raise LinAlgError("not positive definite, even with jitter.")


def pdinv(A):
    """(A^-1 symmetric, L, logdet) as gpy_linalg.py:219-240 (minus the unused Li)."""
    L = jitchol(A)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    Ai, _ = lapack.dpotri(L, lower=1)
    Ai = np.tril(Ai) + np.tril(Ai, -1).T      # symmetrify, gpy_linalg.py:236-238
    return Ai, L, logdet


def dpotrs(L, B):
    """Solve (L L^T) X = B (gpy_linalg.py:139-148)."""
    X, info = lapack.dpotrs(np.asfortranarray(L), B, lower=1)
    return X
