"""Python-3 L-BFGS driver around ``GPRF.llgrad`` - the caller of the hot path.

Mirrors the reference's experiment driver for the GPRF objective:

  do_optimization   gprfopt.py:322-432   objective/gradient callback handed to
                                         scipy.optimize.minimize (L-BFGS-B, ftol 1e-6, maxiter 200):
                                         update_X / update_covs -> llgrad -> + x_prior / cov_prior,
                                         log-theta chain rule, ``log.txt`` and ``step_%05d_X.npy``
  analyze_run       gprfopt.py:453-515   ``results.txt``: step time ll lscale_ratio mad xprior ... and
                                         the final ``trueX`` line; prediction columns (SMSE / MSLL of
                                         gprfopt.py:121-170 through GPRF.train_predictor) with
                                         predict=True, else 0 as without --analyze_full
  do_run            gprfopt.py:525-580   task = x | cov | xcov initialisation
  build_run_name    gprfopt.py:583-597   directory naming of a run

The ``gprf`` object only needs the reference's method surface (update_X, update_covs, llgrad,
cov, n_blocks), so the same driver runs the CUDA-backed ``gprf_b200.GPRF`` and the CPU oracle;
``tests/test_lbfgs_trajectory.py`` checks both against the trajectories the reference itself
logged (tests/golden/gprf_trajectories_golden.json).
"""
import os
import time

import numpy as np
import scipy.optimize

COV_SCALE = 5.0        # gprfopt.py:366 "hack to better condition optimization of cov params"


class OutOfTimeError(Exception):
    pass


def cov_prior(c):
    """Near-uniform Gaussian prior on the log covariance parameters (gprfopt.py:324-331)."""
    mean, std = -1.0, 10.0
    r = (c - mean) / std
    return -.5 * np.sum(r ** 2) - .5 * len(c) * np.log(2 * np.pi * std ** 2), -(c - mean) / std ** 2


def _full_cov(C, sdata, dx):
    """(1,1) lengthscale or (1,4) full theta -> full theta row (gprfopt.py:333-346)."""
    if C.shape[1] == 1:
        FC = np.empty((C.shape[0], 2 + dx))
        FC[:, 0] = sdata.noise_var
        FC[:, 1] = 1.0
        FC[:, 2:] = C
        return FC
    if C.shape[1] == 2 + dx:
        return C
    raise ValueError("unrecognized cov param shape %r" % (C.shape,))


def _collapse_cov_grad(g, C0):
    """Gradient w.r.t. the optimised parametrisation (gprfopt.py:348-356)."""
    if C0.shape[1] == 1:
        return np.sum(g[:, 2:], axis=1, keepdims=True)
    return g


def do_optimization(d, gprf, X0, C0, sdata, method="l-bfgs-b", maxsec=3600, parallel=False, maxiter=200,
                    max_evals=None, save_steps=True, verbose=False, fused=True):
    """gprfopt.py:322-432.  Returns the list of (step, seconds, objective) that is also written to
    ``log.txt``.  ``max_evals`` (not in the reference) stops after that many callback evaluations.
    ``fused``: with a GPRF that offers ``neg_objective`` the location prior (gprfopt.py:172-182), the sum
    with the likelihood gradient and the sign change happen on the device in the evaluation's epilogue
    (same arithmetic, gprf_neg_objective); only (f, g) come back."""
    grad_X, grad_C = X0 is not None, C0 is not None
    fused = bool(fused and grad_X and hasattr(gprf, "neg_objective") and hasattr(sdata, "X_obs")
                 and hasattr(sdata, "obs_std"))
    if fused:
        gprf.set_x_prior(sdata.X_obs, sdata.obs_std)
    x0 = X0.flatten() if grad_X else np.array(())
    c0 = np.log(C0.flatten()) * COV_SCALE if grad_C else np.array(())
    full0 = np.concatenate([x0, c0])
    os.makedirs(d, exist_ok=True)
    log = []
    t0 = time.time()
    dx = sdata.X_obs.shape[1]

    with open(os.path.join(d, "log.txt"), "w") as f_log:
        def neg_llgrad(x):
            if time.time() - t0 > maxsec or (max_evals is not None and len(log) >= max_evals):
                raise OutOfTimeError
            step = len(log)
            xx, xc = x[:len(x0)], x[len(x0):] / COV_SCALE
            if grad_X:
                XX = xx.reshape(X0.shape)
                if not fused:
                    gprf.update_X(XX)
                if save_steps:
                    np.save(os.path.join(d, "step_%05d_X.npy" % step), XX)
            if grad_C:
                C = np.exp(xc.reshape(C0.shape))
                FC = _full_cov(C, sdata, dx)
                gprf.update_covs(FC)
                if save_steps:
                    np.save(os.path.join(d, "step_%05d_cov.npy" % step), FC)
            parts = []
            if fused:
                f_neg, g_neg, gC = gprf.neg_objective(XX, grad_cov=grad_C)
                ll = -f_neg
                parts.append(-g_neg.reshape(-1))
            else:
                ll, gX, gC = gprf.llgrad(local=True, grad_X=grad_X, grad_cov=grad_C, parallel=parallel)
            if grad_X and not fused:
                pl, pg = sdata.x_prior(xx)
                ll += pl
                parts.append(gX.flatten() + pg)
            if grad_C:
                pl, pg = cov_prior(xc)
                ll += pl
                parts.append(((_collapse_cov_grad(np.asarray(gC), C0) * C).flatten() + pg) / COV_SCALE)
            sec = time.time() - t0
            log.append((step, sec, float(ll)))
            f_log.write("%d %.2f %.2f\n" % (step, sec, ll))
            f_log.flush()
            if verbose:
                print("%d %.2f %.2f" % (step, sec, ll))
            return -ll, -np.concatenate(parts)

        try:
            scipy.optimize.minimize(neg_llgrad, full0, jac=True, method=method, bounds=None,
                                    options={"ftol": 1e-6, "maxiter": maxiter})
        except OutOfTimeError:
            pass
        f_log.write("optimization finished after %.fs\n" % (time.time() - t0))
    open(os.path.join(d, "finished"), "w").close()
    return log


def load_log(d):
    """gprfopt.py:435-450."""
    steps, times, lls = [], [], []
    with open(os.path.join(d, "log.txt")) as lf:
        for line in lf:
            f = line.split(" ")
            try:
                s, t, l = int(f[0]), float(f[1]), float(f[2])
            except (ValueError, IndexError):
                continue
            steps.append(s); times.append(t); lls.append(l)
    return np.asarray(steps), np.asarray(times), np.asarray(lls)


def mean_distance(sdata, x):
    """gprfopt.py:76-79."""
    return float(np.mean(np.linalg.norm(x.reshape(sdata.SX.shape) - sdata.SX, axis=1)))


def analyze_run(d, sdata, local_dist=1.0, build_gprf=None, predict=False, predict_kwargs=None):
    """gprfopt.py:453-515.  ``predict`` (the reference's --analyze_full) fills the six prediction
    columns smse_local smse msll_local_block msll_block msll_local_diag msll_diag through
    ``sdata.prediction_error`` (gprfopt.py:121-170); without it they are 0 as in the shipped logs.
    Returns the rows."""
    steps, times, lls = load_log(d)
    rows = []
    pk = predict_kwargs or {}

    def pred_cols(X, FC):
        if not predict:
            return (0., 0., 0., 0., 0., 0.)
        sl, bl, dl = sdata.prediction_error(X=X, cov=FC, local_dist=1.0, **pk)
        s, b, dg = sdata.prediction_error(X=X, cov=FC, local_dist=local_dist, **pk) if local_dist < 1.0 else (sl, bl, dl)
        return (sl, s, bl, b, dl, dg)

    with open(os.path.join(d, "results.txt"), "w") as res:
        for i, step in enumerate(steps):
            try:
                X = np.load(os.path.join(d, "step_%05d_X.npy" % step))
            except IOError:
                X = sdata.SX
            try:
                FC = np.load(os.path.join(d, "step_%05d_cov.npy" % step))
            except IOError:
                FC = None
            mad = mean_distance(sdata, X.flatten())
            c1 = FC[0, 2] / sdata.cov.dfn_params[0] if FC is not None else 0.0      # lscale_error, :90-93
            xp = sdata.x_prior(X.flatten())[0]
            pc = pred_cols(X, FC)
            rows.append((int(step), times[i], lls[i], c1, mad, xp) + (pc if predict else ()))
            res.write("%d %.2f %.2f %.8f %.8f %.8f %.4f %.4f %.4f %.4f %.4f %.4f\n"
                      % ((step, times[i], lls[i], c1, mad, xp) + pc))
        ll1 = -np.inf
        if build_gprf is not None:
            g = build_gprf(X=sdata.SX, local_dist=local_dist)
            try:
                if g.n_blocks > 1:
                    ll1 = g.llgrad()[0]
            except Exception:      # the reference swallows any failure here (gprfopt.py:507-511)
                pass
        res.write("trueX inf %.2f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f\n"
                  % ((ll1, 0.0, 0.0, sdata.x_prior(sdata.SX.flatten())[0]) + pred_cols(sdata.SX, None)))
    return rows


def initial_point(sdata, gprf, task="x", init_seed=-1, init_true=False):
    """(X0, C0) of gprfopt.py:546-573."""
    if task == "x":
        if init_true:
            gprf.update_X(sdata.SX)
            return sdata.SX, None
        return sdata.X_obs, None
    if task == "cov":
        gprf.update_X(sdata.SX)
        if init_seed >= 0:
            np.random.seed(init_seed)
            return None, np.exp(np.random.randn(1, 4) - 1)
        return None, np.array((0.01, 1.0, 0.05, 0.05)).reshape(1, -1)
    if task == "xcov":
        X0 = sdata.X_obs
        if init_seed >= 0:
            np.random.seed(init_seed)
            C0 = np.exp(np.random.randn(1, 1) - 1)
            return X0 + np.random.randn(*X0.shape) * 0.005, C0
        return X0, np.array(gprf.cov.dfn_params[0]).reshape(1, 1)
    raise ValueError("unrecognized task " + task)


def build_run_name(ntrain, ntest, nblocks, lscale, obs_std, local_dist, yd=50, method="l-bfgs-b", task="x",
                   init_seed=-1, noise_var=0.01, seed=0, rpc_blocksize=-1, gplvm_type="gprf", num_inducing=0,
                   init_true=False):
    """gprfopt.py:583-597."""
    return "%d_%d_%s_%.6f_%.6f_%.4f_%d_%s_%s_%d_%s_s%s_%s%d" % (
        ntrain, ntrain + ntest, "%d" % nblocks if rpc_blocksize == -1 else "%06d" % rpc_blocksize, lscale, obs_std,
        local_dist, yd, method, task, -9999 if init_true else init_seed, "%.4f" % noise_var, "%d" % seed,
        gplvm_type, num_inducing)


def do_run(d, sdata, local_dist=1.0, task="x", init_seed=-1, init_true=False, method="l-bfgs-b", maxsec=3600,
           build_kwargs=None, **opt_kwargs):
    """gprfopt.py:525-580 for gplvm_type == "gprf": build, optimise, analyse."""
    gprf = sdata.build_gprf(local_dist=local_dist, **(build_kwargs or {}))
    X0, C0 = initial_point(sdata, gprf, task, init_seed, init_true)
    log = do_optimization(d, gprf, X0, C0, sdata, method=method, maxsec=maxsec, **opt_kwargs)
    rows = analyze_run(d, sdata, local_dist=local_dist,
                       build_gprf=lambda **kw: sdata.build_gprf(**dict(kw, **(build_kwargs or {}))))
    return gprf, log, rows
