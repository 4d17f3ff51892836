"""Data generation at scale on the GPU (synthetic.py:103-137) pinned against
the reference's own result logs.

``gprf_b200.synthetic.sample_y(..., device=0)`` draws Y = jitchol(K + nv I) Z with the evaluation
path's batched Cholesky (one unit holding all n + 500 points).  With it every ``*_gprf0`` run of
gprf_results.tgz whose data came from the reference's DENSE branch (n < 40000: 75 runs, ntrain =
2000 ... 35000) is regenerated from the seed recipe and its logged step-0 and trueX objective
values are reproduced by the CUDA llgrad to the printed precision (+-0.006 on values of 1e5..1e8,
about 1e-9 relative).  Runs with ntrain >= 40000 used CHOLMOD's sparse factorisation, whose sample
depends on that library's fill-reducing permutation: not reproducible (see sample_y_device).
"""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_oracle_golden import _cov_prior_at_init, TOL  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "gprf_results_golden.json")))["runs"]
NTRAINS = sorted(set(r["ntrain"] for r in GOLD if r["ntrain"] < 40000))
# Four xcov logs whose first results.txt row is not the evaluation at X_obs (their other rows and
# their trueX value are consistent with the data; the run lengths 225 / 214 / 13 / 219 rows are outliers
# too - restarted runs).  Every x run and every other xcov run of the same data sets reproduces, as do
# these four runs' own trueX values: profiles/r02_golden_runs_device_sampling.txt.
STEP0_NOT_AT_XOBS = {
    "15000_15500_64_0.048990_0.016330_1.0000_50_l-bfgs-b_xcov_-1_0.0100_s0_gprf0",
    "25000_25500_121_0.037947_0.012649_1.0000_50_l-bfgs-b_xcov_-1_0.0100_s0_gprf0",
    "30000_30500_144_0.034641_0.011547_0.1000_50_l-bfgs-b_xcov_-1_0.0100_s0_gprf0",
    "30000_30500_144_0.034641_0.011547_1.0000_50_l-bfgs-b_xcov_-1_0.0100_s0_gprf0",
}


@pytest.mark.parametrize("ntrain", NTRAINS)
def test_device_sample_reproduces_golden_runs(ntrain):
    from gprf_b200 import grid_centers
    from gprf_b200.synthetic import SampledData
    runs = [r for r in GOLD if r["ntrain"] == ntrain]
    assert runs and all(r["seed"] == 0 and r["n"] == ntrain + 500 and r["yd"] == 50 for r in runs)
    sd = SampledData(noise_var=0.01, n=ntrain + 500, ntrain=ntrain, lscale=6.0 / np.sqrt(ntrain),
                     obs_std=2.0 / np.sqrt(ntrain), yd=50, seed=0, device=0)
    assert abs(sd.lscale - runs[0]["lscale"]) < 1e-6 and abs(sd.obs_std - runs[0]["obs_std"]) < 1e-6
    xp_obs = sd.x_prior(sd.X_obs.flatten())[0]
    checked = 0
    for run in runs:
        xp0 = sd.x_prior(sd.SX.flatten())[0] if run["init_seed"] == -9999 else xp_obs
        assert abs(xp0 - run["step0_xprior"]) < 1e-4
        sd.set_centers(grid_centers(run["nblocks"]))
        if run["init_seed"] == -9999:       # --init_true: starts from SX, blocks recomputed on SX
            gp = sd.build_gprf(local_dist=run["local_dist"])
            gp.update_X(sd.SX)
            ll0 = gp.llgrad()[0] + sd.x_prior(sd.SX.flatten())[0]
        else:
            gp = sd.build_gprf(local_dist=run["local_dist"])
            ll0 = gp.llgrad()[0] + xp_obs
        gp.close()
        if run["task"] == "xcov":
            ll0 += _cov_prior_at_init(sd.cov.dfn_params[0])
        if run["dir"] not in STEP0_NOT_AT_XOBS:
            assert abs(ll0 - run["step0_ll"]) < TOL + 2e-10 * abs(run["step0_ll"]), (run["dir"], ll0, run["step0_ll"])
        if run.get("trueX_ll") is not None:
            gt = sd.build_gprf(X=sd.SX, local_dist=run["local_dist"])
            llt = gt.llgrad()[0]
            gt.close()
            assert abs(llt - run["trueX_ll"]) < TOL + 2e-10 * abs(run["trueX_ll"]), (run["dir"], llt, run["trueX_ll"])
        checked += 1
    assert checked == len(runs)


def test_device_sample_other_families_and_statistics():
    """The device draw is the exact L z of the same covariance for every family: against numpy on a
    small lld + Matern-3/2 set (Cholesky is unique, so L z agrees to rounding)."""
    from gprf_b200 import GPCov
    from gprf_b200.synthetic import sample_y_device
    from oracle import kernels as kern
    from oracle.kernels import GPCov as OCov
    rng = np.random.RandomState(2)
    n = 700
    X = np.column_stack([80 + 2 * rng.rand(n), 30 + 2 * rng.rand(n), 40 * rng.rand(n)])
    th = dict(wfn_params=[1.3], dfn_params=[60.0, 25.0], dfn_str="lld", wfn_str="matern32")
    Z = rng.randn(n, 50)
    Y = sample_y_device(X, GPCov(**th), 0.05, 50, device=0, Z=Z)
    K = kern.kernel_matrix(X, X, OCov(**th)) + 0.05 * np.eye(n)
    want = np.linalg.cholesky(K) @ Z
    assert np.abs(Y - want).max() <= 1e-9 * np.abs(want).max()
