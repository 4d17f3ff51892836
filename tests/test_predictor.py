"""train_predictor / prediction_error (gprf.py:593-672, gprfopt.py:121-170): oracle properties on
CPU, CUDA-vs-oracle parity on the GPU."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def small_data(cls_mod):
    sd = cls_mod.SampledData(noise_var=0.01, n=700, ntrain=600, lscale=0.15, obs_std=0.01, yd=4, seed=2)
    grid = [(x, y) for x in (1 / 6., 3 / 6., 5 / 6.) for y in (1 / 6., 3 / 6., 5 / 6.)]
    sd.set_centers(np.array(grid))
    return sd


def test_oracle_single_block_is_exact_gp():
    """With one block the committee is a single GP: the fused prediction must be its posterior."""
    from oracle import synthetic as osyn
    from oracle import kernels as kern
    sd = osyn.SampledData(noise_var=0.01, n=160, ntrain=120, lscale=0.3, obs_std=0.01, yd=3, seed=1)
    sd.set_centers(np.array([[0.5, 0.5]]))
    g = sd.build_gprf(local_dist=1.0)
    p = g.train_predictor()
    Xs = sd.Xtest[:7]
    m, c = p(Xs, test_noise_var=0.0)
    K = kern.kernel_matrix(g.X, g.X, g.cov) + 0.01 * np.eye(120)
    Ks = kern.kernel_matrix(Xs, g.X, g.cov)
    want_m = Ks.dot(np.linalg.solve(K, g.Y))
    want_c = kern.kernel_matrix(Xs, Xs, g.cov) - Ks.dot(np.linalg.solve(K, Ks.T))
    assert np.allclose(m, want_m, rtol=1e-6, atol=1e-8)
    assert np.allclose(c, want_c, rtol=1e-5, atol=1e-9)


def test_oracle_prediction_error_sane():
    from oracle import synthetic as osyn
    sd = small_data(osyn)
    smse_l, msll_l, msll_dl = sd.prediction_error(X=sd.SX, local_dist=1.0)
    smse, msll, msll_d = sd.prediction_error(X=sd.SX, local_dist=0.4)
    assert 0 < smse < smse_l * 1.05 < 0.6          # neighbours help (or at least do not hurt) and beat the mean
    assert msll > 0 and msll_l > 0                  # better than the trivial Gaussian baseline
    noisy = sd.prediction_error(X=sd.SX + 0.05 * np.random.RandomState(0).randn(*sd.SX.shape), local_dist=0.4)
    assert noisy[0] > smse                          # wrong locations predict worse


def test_oracle_test_cov_enters_the_prior_only():
    """gprf.py:599-605,621 vs :649-654: test_cov builds the prior covariance; the per-block messages
    (Kstar, Kss) keep the training covariance.  With one block and the prior's extra precision removed,
    the posterior mean is therefore independent of test_cov only through that prior term."""
    from oracle import synthetic as osyn
    from oracle.kernels import GPCov as OCov
    sd = osyn.SampledData(noise_var=0.01, n=160, ntrain=120, lscale=0.3, obs_std=0.01, yd=3, seed=1)
    sd.set_centers(np.array([[0.5, 0.5]]))
    g = sd.build_gprf(local_dist=1.0)
    other = OCov(wfn_params=[2.5], dfn_params=[0.11, 0.11], dfn_str="euclidean", wfn_str="se")
    Xs = sd.Xtest[:6]
    m0, c0 = g.train_predictor()(Xs)
    m1, c1 = g.train_predictor(test_cov=other)(Xs)
    # one source block: final precision = inv(prior_cov) + inv(cov) - inv(Kss); only the first term changes
    from oracle import kernels as kern
    d_prec = np.linalg.inv(c1) - np.linalg.inv(c0)
    want = np.linalg.inv(kern.kernel_matrix(Xs, Xs, other)) - np.linalg.inv(kern.kernel_matrix(Xs, Xs, g.cov))
    assert np.allclose(d_prec, want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
    assert not np.allclose(m0, m1)


@pytest.mark.gpu
def test_predictor_test_cov_cuda_matches_oracle():
    from oracle import synthetic as osyn
    from gprf_b200 import synthetic as psyn
    from gprf_b200 import GPCov
    from oracle.kernels import GPCov as OCov
    so, sp = small_data(osyn), small_data(psyn)
    go, gp = so.build_gprf(local_dist=0.4), sp.build_gprf(local_dist=0.4)
    th = dict(wfn_params=[1.7], dfn_params=[0.2, 0.25], dfn_str="euclidean", wfn_str="se")
    po, pp = go.train_predictor(test_cov=OCov(**th)), gp.train_predictor(test_cov=GPCov(**th))
    Xs = so.Xtest[:9]
    for tnv in (0.0, 0.01):
        mo, co = po(Xs, test_noise_var=tnv)
        mp, cp = pp(Xs, test_noise_var=tnv)
        assert np.abs(mo - mp).max() <= 1e-7 * np.abs(mo).max()
        assert np.abs(co - cp).max() <= 1e-7 * np.abs(co).max()


@pytest.mark.gpu
def test_predictor_cuda_matches_oracle(tmp_path):
    from oracle import synthetic as osyn
    from gprf_b200 import synthetic as psyn
    from gprf_b200 import gprfopt
    so, sp = small_data(osyn), small_data(psyn)
    assert np.array_equal(so.SX, sp.SX) and np.array_equal(so.Ytest, sp.Ytest)
    go, gp = so.build_gprf(local_dist=0.4), sp.build_gprf(local_dist=0.4)
    Ko, Ao = [np.linalg.inv(go.kernel(go.X[i])) for i in go.block_idxs], None
    Kp, Ap = gp.block_precisions()
    for a, b in zip(Ko, Kp):
        assert np.abs(a - b).max() <= 1e-8 * np.abs(a).max()
    po, pp = go.train_predictor(), gp.train_predictor()
    Xs = so.Xtest[:9]
    for tnv in (0.0, 0.01):
        mo, co = po(Xs, test_noise_var=tnv)
        mp, cp = pp(Xs, test_noise_var=tnv)
        assert np.abs(mo - mp).max() <= 1e-7 * np.abs(mo).max()
        assert np.abs(co - cp).max() <= 1e-7 * np.abs(co).max()
    eo = so.prediction_error(X=so.SX, local_dist=0.4)
    ep = sp.prediction_error(X=sp.SX, local_dist=0.4)
    assert np.allclose(eo, ep, rtol=1e-6, atol=1e-8), (eo, ep)
    # the driver's --analyze_full columns
    d = str(tmp_path / "run")
    g2 = sp.build_gprf(local_dist=0.4)
    gprfopt.do_optimization(d, g2, sp.X_obs, None, sp, max_evals=2, save_steps=True)
    rows = gprfopt.analyze_run(d, sp, local_dist=0.4, predict=True)
    assert len(rows) == 2 and len(rows[0]) == 12 and 0 < rows[0][7] < 1
    last = open(os.path.join(d, "results.txt")).read().splitlines()[-1].split()
    assert last[0] == "trueX" and abs(float(last[7]) - ep[0]) < 1e-4
