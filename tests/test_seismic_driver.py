"""The py3 mirror of the reference's seismic L-BFGS driver (gprf_b200/run_seismic.py, following
run_seismic.py:69-215, 235-289, 340-415), driven by the CPU oracle here and by the CUDA GPRF in
the gpu-marked test, which must walk the same optimisation trajectory."""
import doctest
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def catalogue(n=360, dy=6, seed=3):
    """A compact cluster of events (lon, lat, depth km) so that neighbouring PD-tree blocks correlate."""
    rng = np.random.RandomState(seed)
    X = np.column_stack([80 + 0.5 * rng.randn(n), 35 + 0.4 * rng.randn(n), np.clip(rng.exponential(30.0, n), 0, 700)])
    return X, rng.randn(n, dy)


def oracle_problem(tmp, task="xcov", threshold=0.6):
    from gprf_b200 import run_seismic as rs
    from oracle.blocking import pdtree_cluster
    from oracle.gprf_oracle import OracleGPRF
    from oracle.kernels import GPCov
    X, Y = catalogue()
    cov = GPCov(wfn_params=[1.0], dfn_params=[40.0, 40.0], dfn_str="lld", wfn_str="matern32")
    return rs.setup_seismic(X, Y, cov, obs_std=2.0, seed=0, block_size=70, threshold=threshold, task=task,
                            cache_dir=str(tmp), gprf_cls=OracleGPRF, pdtree_fn=pdtree_cluster)


def test_great_circle_doctests_and_priors():
    from gprf_b200 import run_seismic as rs
    assert doctest.testmod(rs).failed == 0
    assert abs(rs.dist_km((10, 0), (20, 0)) - np.radians(10) * 6371.0) < 1e-9
    assert abs(rs.dist_lld((10, 0, 5.0), (10, 0, 9.0)) - 4.0) < 1e-12
    c = np.array([-2.3, 0.0, 3.6, 3.6])
    ll, g = rs.cov_prior(c)
    assert abs(ll + 0.5 * 4 * np.log(2 * np.pi * 1.5 ** 2)) < 1e-12 and np.allclose(g, 0)
    c2 = c.copy()
    c2[2] = 5.01                                           # large-lengthscale penalty, run_seismic.py:83-87
    ll2, g2 = rs.cov_prior(c2)
    pen = np.exp(70 * 0.01)
    assert abs(ll2 - (ll - 0.5 * ((5.01 - 3.6) / 1.5) ** 2 - pen)) < 1e-9
    assert abs(g2[2] - (-(5.01 - 3.6) / 1.5 ** 2 - 70 * pen)) < 1e-9
    FC = rs.clamp_cov(np.array([[50.0, 3.0, 0.2, 5000.0]]))
    assert FC.tolist() == [[10.0, 1.0, 1.0, 999.0]]
    # x_prior: gradient is the derivative of the value
    means = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    xp = rs.make_x_prior(means, np.array([0.02, 0.02, 2.0]))
    X = means + 0.01
    l0, g0 = xp(X)
    Xe = X.copy()
    Xe[1, 2] += 1e-6
    assert abs((xp(Xe)[0] - l0) / 1e-6 - g0[1, 2]) < 1e-4 * abs(g0[1, 2])


def test_seismic_driver_oracle(tmp_path):
    from gprf_b200 import run_seismic as rs
    prob = oracle_problem(tmp_path)
    assert os.path.exists(prob.neighbor_file)                  # neighbors_%d_%d_%.3f_%.3f.npy, (E, 2) ints
    saved = np.load(prob.neighbor_file)
    assert saved.ndim == 2 and saved.shape[1] == 2 and len(saved) == len(prob.gprf.neighbors) > 0
    assert os.path.basename(prob.neighbor_file) == "neighbors_360_70_0.600_2.000.npy"
    again = oracle_problem(tmp_path)                           # second run reads the cache
    assert [tuple(e) for e in again.gprf.neighbors] == [tuple(e) for e in prob.gprf.neighbors]
    X0_before = prob.X0.copy()
    d = str(tmp_path / "run")
    log = rs.do_optimization(d, prob.gprf, prob.X0, prob.C0, rs.cov_prior, prob.x_prior, max_evals=4)
    assert np.array_equal(prob.X0, X0_before)                  # caller's X0 is not rescaled in place
    assert len(log) == 4 and all(np.isfinite(r[2]) for r in log)
    lines = open(os.path.join(d, "log.txt")).read().splitlines()
    assert lines[0].split()[0] == "0" and lines[-1].startswith("optimization finished after")
    assert len(open(os.path.join(d, "covs.txt")).read().splitlines()) >= 4
    assert os.path.exists(os.path.join(d, "finished")) and os.path.exists(os.path.join(d, "step_00003_X.npy"))
    FC = np.load(os.path.join(d, "step_00000_cov.npy"))
    assert FC.shape == (1, 4) and FC[0, 1] == 1.0
    X1 = np.load(os.path.join(d, "step_00001_X.npy"))
    assert X1.shape == prob.X_true.shape and abs(np.median(X1[:, 2]) - np.median(prob.X_true[:, 2])) < 10.0
    rows = rs.analyze_run_result(d, prob)
    assert len(rows) == 5 and rows[-1].startswith("true X ll ")
    assert len(rows[0].split()) == 6
    # task = x and task = cov (the reference crashes on the latter)
    for task in ("x", "cov"):
        p2 = oracle_problem(tmp_path, task=task)
        lg = rs.do_optimization(str(tmp_path / task), p2.gprf, p2.X0, p2.C0, rs.cov_prior, p2.x_prior, max_evals=2,
                                save_steps=False)
        assert len(lg) == 2 and np.isfinite(lg[0][2])


def test_seismic_first_objective_is_llgrad_plus_priors(tmp_path):
    from gprf_b200 import run_seismic as rs
    prob = oracle_problem(tmp_path)
    ll, gX, gC = prob.gprf.llgrad(grad_X=True, grad_cov=True)
    want = ll + prob.x_prior(prob.X0)[0] + rs.cov_prior(np.log(prob.C0.flatten()))[0]
    log = rs.do_optimization(str(tmp_path / "r"), prob.gprf, prob.X0, prob.C0, rs.cov_prior, prob.x_prior,
                             max_evals=1, save_steps=False)
    assert abs(log[0][2] - want) <= 1e-9 * abs(want)


@pytest.mark.gpu
def test_seismic_driver_cuda_matches_oracle(tmp_path):
    """Same catalogue, same driver: the CUDA GPRF and the oracle produce the same objective at every
    one of the first evaluations of the L-BFGS run (edges, blocks, depth rescaling, priors, clamps)."""
    from gprf_b200 import GPCov, run_seismic as rs
    po = oracle_problem(tmp_path / "o")
    X, Y = catalogue()
    cov = GPCov([1.0], [40.0, 40.0], "lld", "matern32")
    os.makedirs(str(tmp_path / "g"), exist_ok=True)
    pg = rs.setup_seismic(X, Y, cov, obs_std=2.0, seed=0, block_size=70, threshold=0.6, task="xcov",
                          cache_dir=str(tmp_path / "g"))
    assert [tuple(e) for e in pg.gprf.neighbors] == [tuple(e) for e in po.gprf.neighbors]
    assert np.array_equal(np.load(pg.neighbor_file), np.load(po.neighbor_file))
    lo = rs.do_optimization(str(tmp_path / "o" / "run"), po.gprf, po.X0, po.C0, rs.cov_prior, po.x_prior, max_evals=6)
    lg = rs.do_optimization(str(tmp_path / "g" / "run"), pg.gprf, pg.X0, pg.C0, rs.cov_prior, pg.x_prior, max_evals=6)
    assert len(lo) == len(lg) == 6
    for a, b in zip(lo, lg):
        assert abs(a[2] - b[2]) <= 1e-8 * abs(a[2]), (a, b)
    ro = rs.analyze_run_result(str(tmp_path / "o" / "run"), po)
    rg = rs.analyze_run_result(str(tmp_path / "g" / "run"), pg)
    assert abs(float(ro[-1].split()[-1]) - float(rg[-1].split()[-1])) <= 0.011
