"""Covariance-function arithmetic of treegp's ``VectorTree`` (oracle side).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

treegp (pinned by the reference's README.md:4 to commit
a0aa7ae65a4b9144a499016bbf0ccaf0c611cc0d) is *not* vendored in
``/root/reference``; what is restated here are the published definitions as
used at the reference's call sites:

* ``VectorTree.kernel_matrix(X1, X2, distance_only)``   gprf.py:339,342,373
* ``VectorTree.kernel_deriv_wrt_xi_row(X, p, i, out)``  gprf.py:353
* ``VectorTree.kernel_deriv_wrt_i(X1, X2, i, 1, dists)``gprf.py:374

distance functions ``dfn_str``:
  "euclidean"  r = sqrt(sum_i (dx_i / l_i)^2)
  "lld"        x = (lon deg, lat deg, depth km); l = (l_horiz km, l_depth km)
               d = great-circle km (run_seismic.py:19-63, R = 6371.0),
               r = sqrt((d/l0)^2 + (dz/l1)^2)   (cf. dist_lld, run_seismic.py:230-233)
weight functions ``wfn_str`` (single parameter sigma^2):
  "se"         w(r) = s2 * exp(-r^2)           (no 1/2: gprfopt.py:238-239
                                                 feeds GPy lengthscale sqrt(.5)*l)
  "matern32"   w(r) = s2 * (1 + sqrt3 r) exp(-sqrt3 r)

The SE/euclidean pair is pinned by the golden objective values; the
Matern/lld pair is unpinned (see package docstring).
"""
import numpy as np

EARTH_RADIUS_KM = 6371.0          # run_seismic.py:43
DEG = np.pi / 180.0
SQRT3 = np.sqrt(3.0)


class GPCov(object):
    """Plain record standing in for ``treegp.gp.GPCov`` (gprf.py:163)."""

    def __init__(self, wfn_params, dfn_params, dfn_str="euclidean", wfn_str="se"):
        self.wfn_params = np.array(wfn_params, dtype=float).reshape(-1)
        self.dfn_params = np.array(dfn_params, dtype=float).reshape(-1)
        self.dfn_str = dfn_str
        self.wfn_str = wfn_str


def great_circle_km(lon1, lat1, lon2, lat2):
    """Haversine distance, following run_seismic.py:19-63 (broadcasts)."""
    p1 = lat1 * DEG
    p2 = lat2 * DEG
    sp = np.sin((p1 - p2) / 2.0)
    sl = np.sin((lon1 * DEG - lon2 * DEG) / 2.0)
    h = sp * sp + np.cos(p1) * np.cos(p2) * sl * sl
    return 2.0 * EARTH_RADIUS_KM * np.arcsin(np.sqrt(h))


def _components(X1, X2, dfn_str):
    """Unscaled distance components c_t[p, q] (one per lengthscale)."""
    if dfn_str == "euclidean":
        return [X1[:, None, t] - X2[None, :, t] for t in range(X1.shape[1])]
    if dfn_str == "lld":
        d = great_circle_km(X1[:, None, 0], X1[:, None, 1], X2[None, :, 0], X2[None, :, 1])
        return [d, X1[:, None, 2] - X2[None, :, 2]]
    raise ValueError("unknown distance function %r" % dfn_str)


def scaled_distance(X1, X2, cov):
    comps = _components(X1, X2, cov.dfn_str)
    r2 = 0.0
    for c, l in zip(comps, cov.dfn_params):
        r2 = r2 + (c / l) ** 2
    return np.sqrt(r2)


def weight(r, cov):
    s2 = cov.wfn_params[0]
    if cov.wfn_str == "se":
        return s2 * np.exp(-r * r)
    if cov.wfn_str == "matern32":
        return s2 * (1.0 + SQRT3 * r) * np.exp(-SQRT3 * r)
    raise ValueError("unknown weight function %r" % cov.wfn_str)


def weight_deriv(r, cov):
    """dw/dr."""
    s2 = cov.wfn_params[0]
    if cov.wfn_str == "se":
        return -2.0 * r * s2 * np.exp(-r * r)
    if cov.wfn_str == "matern32":
        return -3.0 * s2 * r * np.exp(-SQRT3 * r)
    raise ValueError("unknown weight function %r" % cov.wfn_str)


def kernel_matrix(X1, X2, cov, distance_only=False):
    """``VectorTree.kernel_matrix`` (gprf.py:339,342,373): no noise term."""
    r = scaled_distance(X1, X2, cov)
    if distance_only:
        return r
    return weight(r, cov)


def _dr_dx1(X1, X2, cov, r):
    """d r[p,q] / d X1[p,i]  -> list over i of (n1, n2) arrays; 0 where r == 0."""
    l = cov.dfn_params
    with np.errstate(divide="ignore", invalid="ignore"):
        if cov.dfn_str == "euclidean":
            out = [(X1[:, None, i] - X2[None, :, i]) / (l[i] ** 2 * r) for i in range(X1.shape[1])]
        else:
            lam1, phi1 = X1[:, None, 0] * DEG, X1[:, None, 1] * DEG
            lam2, phi2 = X2[None, :, 0] * DEG, X2[None, :, 1] * DEG
            hp = (phi1 - phi2) / 2.0
            hl = (lam1 - lam2) / 2.0
            sp, cp = np.sin(hp), np.cos(hp)
            sl, cl = np.sin(hl), np.cos(hl)
            c1, c2 = np.cos(phi1), np.cos(phi2)
            h = sp * sp + c1 * c2 * sl * sl
            d = 2.0 * EARTH_RADIUS_KM * np.arcsin(np.sqrt(h))
            dd_dh = EARTH_RADIUS_KM / np.sqrt(h * (1.0 - h))
            dh_dlon = c1 * c2 * sl * cl * DEG
            dh_dlat = (sp * cp - np.sin(phi1) * c2 * sl * sl) * DEG
            pref = d * dd_dh / (l[0] ** 2 * r)
            out = [pref * dh_dlon, pref * dh_dlat,
                   (X1[:, None, 2] - X2[None, :, 2]) / (l[1] ** 2 * r)]
    return [np.where(np.isfinite(o), o, 0.0) for o in out]


def kernel_deriv_wrt_xi_rows(X, cov):
    """All rows of ``kernel_deriv_wrt_xi_row`` at once.

    Returns dK with dK[i][p, q] = d k(x_p, x_q) / d x_{p,i}, diagonal zeroed
    (gprf.py:353-354).
    """
    r = scaled_distance(X, X, cov)
    wp = weight_deriv(r, cov)
    out = []
    for drdx in _dr_dx1(X, X, cov, r):
        m = wp * drdx
        np.fill_diagonal(m, 0.0)
        out.append(m)
    return out


def kernel_deriv_wrt_xi_row(X, p, i, cov):
    """One row: d k(x_p, x_q)/d x_{p,i} for all q (gprf.py:345-355)."""
    Xp = X[p:p + 1]
    r = scaled_distance(Xp, X, cov)
    row = (weight_deriv(r, cov) * _dr_dx1(Xp, X, cov, r)[i])[0]
    row[p] = 0.0
    return row


def kernel_deriv_wrt_i(X1, X2, t, cov):
    """dK/d l_t (``kernel_deriv_wrt_i``, gprf.py:374): w'(r) * dr/dl_t."""
    comps = _components(X1, X2, cov.dfn_str)
    l = cov.dfn_params
    r = scaled_distance(X1, X2, cov)
    with np.errstate(divide="ignore", invalid="ignore"):
        drdl = -(comps[t] ** 2) / (l[t] ** 3 * r)
    drdl = np.where(np.isfinite(drdl), drdl, 0.0)
    return weight_deriv(r, cov) * drdl
