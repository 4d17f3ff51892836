"""Python-3 mirror of the reference's seismic driver around ``GPRF.llgrad`` (BASELINE configs[3]).

  dist_deg / dist_km / dist_lld   run_seismic.py:19-63, 230-233   great-circle helpers (doctests kept)
  cov_prior                       run_seismic.py:69-89            Gaussian prior on log theta + large-lengthscale penalty
  make_x_prior                    run_seismic.py:357-371          Gaussian prior around the observed locations
  setup_seismic                   run_seismic.py:340-415          noise model, PD-tree blocks, neighbour cache
                                                                  ``neighbors_%d_%d_%.3f_%.3f.npy``, GPRF construction
  do_optimization                 run_seismic.py:92-215           L-BFGS-B callback: depth rescaling (x100), log-theta
                                                                  transform with the clamps of :134-146, gradient
                                                                  clipping :171-173, ``log.txt`` / ``covs.txt`` /
                                                                  ``step_%05d_{X,cov}.npy`` / ``finished``
  analyze_run_result              run_seismic.py:235-289          ``results.txt``

The catalogue itself (``sorted_isc.npy``, run_seismic.py:291-307) is not part of the reference
checkout, so ``setup_seismic`` takes the (lon, lat, depth) array and Y as arguments;
``bench.make_workload("cfg4")`` builds a synthetic catalogue of the same kind.

The ``gprf`` object only needs the reference's method surface, so the same driver runs the
CUDA-backed ``gprf_b200.GPRF`` and the CPU oracle (``gprf_cls`` argument);
``tests/test_seismic_driver.py`` checks that both walk the same optimisation trajectory.

Two deliberate deviations, both of which only remove crashes of the reference:
``do_optimization`` works on a copy of X0 (the reference divides the caller's array by 100 in
place, run_seismic.py:97) and ``task='cov'`` (X0 is None) works (the reference dereferences
None at :97 and indexes the empty gX at :157 - SURVEY.md section 8c).
"""
import os
import time

import numpy as np
import scipy.optimize

AVG_EARTH_RADIUS_KM = 6371.0
DEPTH_SCALE = 100.0     # run_seismic.py:96


class OutOfTimeError(Exception):
    pass


def dist_deg(loc1, loc2):
    """Great-circle distance in degrees between (lon, lat) pairs (run_seismic.py:19-49).

    >>> int(dist_deg((10,0), (20, 0)))
    10
    >>> int(dist_deg((10,0), (10, 45)))
    45
    >>> int(dist_deg((-78, -12), (-10.25, 52)))
    86
    >>> bool(dist_deg((132.86521, -0.45606493), (132.86521, -0.45606493)) < 1e-4)
    True
    """
    lon1, lat1 = loc1
    lon2, lat2 = loc2
    rlon1, rlat1, rlon2, rlat2 = np.radians(lon1), np.radians(lat1), np.radians(lon2), np.radians(lat2)
    h = np.sin((rlat1 - rlat2) / 2.0) ** 2 + np.cos(rlat1) * np.cos(rlat2) * np.sin((rlon1 - rlon2) / 2.0) ** 2
    return np.degrees(2 * np.arcsin(np.sqrt(h)))


def dist_km(loc1, loc2):
    """run_seismic.py:51-63."""
    return np.radians(dist_deg(loc1, loc2)) * AVG_EARTH_RADIUS_KM


def dist_lld(x1, x2):
    """run_seismic.py:230-233: sqrt(horizontal km^2 + depth km^2)."""
    d1 = dist_km((x1[0], x1[1]), (x2[0], x2[1]))
    d2 = x1[2] - x2[2]
    return np.sqrt(d1 ** 2 + d2 ** 2)


def cov_prior(c):
    """run_seismic.py:69-89 (c = log theta, 4 entries)."""
    means = np.array((-2.3, 0.0, 3.6, 3.6))
    std = 1.5
    r = (c - means) / std
    ll = -.5 * np.sum(r ** 2) - .5 * len(c) * np.log(2 * np.pi * std ** 2)
    lderiv = (-(c - means) / (std ** 2)).reshape((-1,))
    c = c.reshape((-1,))
    if c[2] > 5:
        penalty = np.exp(70 * (c[2] - 5))
        ll -= penalty
        lderiv[2] -= 70 * np.exp(70 * (c[2] - 5))
    return ll, lderiv


def make_x_prior(means, prior_std):
    """run_seismic.py:363-371."""
    def x_prior(X):
        r = (X - means) / prior_std
        r2 = r / prior_std
        n = X.shape[0]
        ll = -.5 * np.sum(r.flatten() ** 2) - .5 * n * (3 * np.log(2 * np.pi) + np.sum(np.log(prior_std ** 2)))
        return ll, -r2.reshape(X.shape)
    x_prior.means = np.asarray(means, dtype=np.float64)          # lets the driver move the prior to the device
    x_prior.prior_std = np.asarray(prior_std, dtype=np.float64)
    return x_prior


def neighbor_cache_name(n, block_size, threshold, obs_std):
    """run_seismic.py:377."""
    return "neighbors_%d_%d_%.3f_%.3f.npy" % (n, block_size, threshold, obs_std)


def clamp_cov(FC):
    """The clamps the reference applies to exp(log theta) before update_covs (run_seismic.py:134-146)."""
    FC = np.array(FC, dtype=np.float64)
    FC[0, 1] = 1.0                      # don't learn sv
    FC[0, 0] = min(FC[0, 0], 10.0)
    FC[0, 2] = min(max(FC[0, 2], 1.0), 999.0)
    FC[0, 3] = min(max(FC[0, 3], 1.0), 999.0)
    return FC


class SeismicProblem(object):
    """What run_seismic.main builds between :340 and :415."""

    def __init__(self, gprf, X_true, X0, C0, cov_true, x_prior, lscale_true, neighbor_file):
        self.gprf, self.X_true, self.X0, self.C0 = gprf, X_true, X0, C0
        self.cov_true, self.x_prior, self.lscale_true, self.neighbor_file = cov_true, x_prior, lscale_true, neighbor_file


def setup_seismic(X_true, SY, cov, obs_std, seed=0, block_size=300, threshold=1.0, task="xcov", cache_dir=".",
                  gprf_cls=None, pdtree_fn=None, init_cov=None, init_x=None, **gprf_kwargs):
    """run_seismic.py:340-415 from the catalogue rows onward.  ``cov`` is the generating GPCov
    (lld / matern32 in the reference, :299-301); noise variance 0.1 (:344)."""
    if gprf_cls is None:
        from .gprf import GPRF as gprf_cls
    if pdtree_fn is None:
        from .blocking import pdtree_cluster as pdtree_fn
    X_true = np.ascontiguousarray(X_true, dtype=np.float64)
    cov_true = np.array([0.1, cov.wfn_params[0], cov.dfn_params[0], cov.dfn_params[1]]).reshape((1, -1))
    np.random.seed(seed)
    prior_std = obs_std * np.array([.01, .01, 1.])
    noise = np.random.randn(*X_true.shape) * prior_std
    means = X_true + noise
    X0 = means.copy()
    x_prior = make_x_prior(means, prior_std)
    n = X0.shape[0]
    cluster_idxs, reblock = pdtree_fn(X0, blocksize=block_size)
    os.makedirs(cache_dir, exist_ok=True)
    fname = os.path.join(cache_dir, neighbor_cache_name(n, block_size, threshold, obs_std))
    if threshold == 1.0:
        neighbors = []
    else:
        try:
            neighbors = [tuple(int(v) for v in row) for row in np.load(fname)]
        except (IOError, OSError, ValueError):
            neighbors = None
    C0 = cov_true.copy() if init_cov is None else np.array(init_cov, dtype=np.float64)
    if init_x is not None:
        X0 = np.array(init_x, dtype=np.float64)
    gprf = gprf_cls(X0, SY, reblock, cov, cov_true[0, 0], neighbor_threshold=threshold, block_idxs=cluster_idxs,
                    neighbors=neighbors, **gprf_kwargs)
    if neighbors is None:
        np.save(fname, np.asarray(gprf.neighbors, dtype=np.int64).reshape(-1, 2))
    if task == "x":
        C0 = None
    elif task == "cov":
        X0 = None
    return SeismicProblem(gprf, X_true, X0, C0, cov_true, x_prior, cov.dfn_params[0], fname)


def do_optimization(d, gprf, X0, C0, cov_prior, x_prior, maxsec=3600, parallel=False, sparse=False, max_evals=None,
                    maxiter=None, save_steps=True, verbose=False):
    """run_seismic.py:92-215.  Returns the (step, seconds, objective) rows also written to ``log.txt``.
    ``max_evals`` / ``maxiter`` / ``save_steps`` are additions for bounded test and bench runs.
    When ``x_prior`` comes from ``make_x_prior`` and the GPRF offers ``neg_objective``, the prior, the
    depth rescaling of the gradient and the sign change run on the device (gprf_neg_objective)."""
    gradX, gradC = X0 is not None, C0 is not None
    fused = bool(gradX and not sparse and hasattr(gprf, "neg_objective") and hasattr(x_prior, "means"))
    if fused:
        dxs = np.shape(X0)[1]
        gs = np.ones(dxs)
        gs[2] = DEPTH_SCALE
        gprf.set_x_prior(x_prior.means, np.broadcast_to(x_prior.prior_std, (dxs,)), grad_scale=gs)
    if gradX:
        X0 = np.array(X0, dtype=np.float64)
        X0[:, 2] /= DEPTH_SCALE
        x0 = X0.flatten()
    else:
        x0 = np.array(())
        X_fixed = np.array(gprf.X, dtype=np.float64)
    c0 = np.log(np.asarray(C0, dtype=np.float64).flatten()) if gradC else np.array(())
    full0 = np.concatenate([x0, c0])
    os.makedirs(d, exist_ok=True)
    log = []
    t0 = time.time()
    kw = {"sparse": sparse} if sparse else {}

    with open(os.path.join(d, "log.txt"), "w") as f_log, open(os.path.join(d, "covs.txt"), "w") as covf:
        def lgpllgrad(x):
            if max_evals is not None and len(log) >= max_evals:
                raise OutOfTimeError
            step = len(log)
            xx, xc = x[:len(x0)], x[len(x0):]
            if gradX:
                XX = xx.reshape(X0.shape).copy()
                XX[:, 2] *= DEPTH_SCALE
                if not fused:
                    gprf.update_X(XX)
                if save_steps:
                    np.save(os.path.join(d, "step_%05d_X.npy" % step), XX)
            else:
                XX = X_fixed
            if gradC:
                FC = clamp_cov(np.exp(xc.reshape(np.shape(C0))))
                gprf.update_covs(FC)
                if save_steps:
                    np.save(os.path.join(d, "step_%05d_cov.npy" % step), FC)
            try:
                if fused:
                    f_neg, g_neg, gC = gprf.neg_objective(XX, grad_cov=gradC)
                    ll = -f_neg
                else:
                    ll, gX, gC = gprf.llgrad(local=True, grad_X=gradX, grad_cov=gradC, parallel=parallel, **kw)
            except Exception as e:            # run_seismic.py:153-155: any failure is a huge objective
                if verbose:
                    print("fail", e)
                return 1e10, np.random.randn(*x.shape)
            parts = []
            if fused:
                parts.append(-g_neg.reshape(-1))
            elif gradX:
                gX = np.array(gX)
                gX[:, 2] *= DEPTH_SCALE
                prior_ll, prior_grad = x_prior(XX)
                prior_grad = np.array(prior_grad)
                prior_grad[:, 2] *= DEPTH_SCALE
                ll += prior_ll
                parts.append(gX.flatten() + prior_grad.flatten())
            if gradC:
                prior_ll, prior_grad = cov_prior(xc)
                ll += prior_ll
                gC = (np.asarray(gC) * FC).flatten() + prior_grad
                gC[1] = 0.0                   # don't learn sv
                max_grad = np.max(np.abs(gC[2:]))
                if max_grad > 10:
                    gC[2:] *= 2. / (1 + max_grad / 10.)
                parts.append(gC.flatten())
            sec = time.time() - t0
            log.append((step, sec, float(ll)))
            f_log.write("%d %.2f %.2f\n" % (step, sec, ll))
            f_log.flush()
            if gradC:
                covf.write("%d %s\n" % (step, FC))
                covf.flush()
            if verbose:
                print("%d %.2f %.2f" % (step, sec, ll))
            if time.time() - t0 > maxsec:
                raise OutOfTimeError
            return -ll, -np.concatenate(parts)

        try:
            opts = {} if maxiter is None else {"maxiter": maxiter}
            scipy.optimize.minimize(lgpllgrad, full0, jac=True, method="l-bfgs-b", bounds=None, options=opts)
        except OutOfTimeError:
            pass
        f_log.write("optimization finished after %.fs\n" % (time.time() - t0))
    open(os.path.join(d, "finished"), "w").close()
    return log


def load_log(d):
    """gprfopt.py:435-450 (imported by run_seismic.py:13)."""
    from .gprfopt import load_log as _ll
    return _ll(d)


def analyze_run_result(d, prob, max_mad_points=None):
    """run_seismic.py:235-289: ``results.txt`` rows ``step time ll lscale_ratio mean_dist median_dist``
    and the final ``true X ll`` line.  ``max_mad_points`` bounds the per-point python loop."""
    steps, times, lls = load_log(d)
    X_true = prob.X_true
    sel = slice(None) if max_mad_points is None else slice(0, max_mad_points)

    def mad(X1, X2):
        dists = [dist_lld(a, b) for a, b in zip(X1[sel], X2[sel])]
        return np.mean(dists), np.median(dists)

    rows = []
    with open(os.path.join(d, "results.txt"), "w") as results:
        for i, step in enumerate(steps):
            try:
                X = np.load(os.path.join(d, "step_%05d_X.npy" % step))
            except IOError:
                X = X_true
            try:
                FC = np.load(os.path.join(d, "step_%05d_cov.npy" % step))
            except IOError:
                FC = None
            c1 = FC[0, 2] / prob.lscale_true if FC is not None else 1.0
            l1, l2 = mad(X_true, X)
            s = "%d %.2f %.2f %.8f %.8f %.8f" % (step, times[i], lls[i], c1, l1, l2)
            rows.append(s)
            results.write(s + "\n")
        prob.gprf.update_X(X_true)
        prob.gprf.update_covs(prob.cov_true)
        lltrue = prob.gprf.llgrad(grad_X=False, grad_cov=False)[0]
        s = "true X ll %.2f" % (lltrue + prob.x_prior(X_true)[0])
        rows.append(s)
        results.write(s + "\n")
    return rows
