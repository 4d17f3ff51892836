"""Pin the CPU oracle against the reference's shipped golden objective values.

Golden source: /root/reference/gprf_results.tgz (see tests/golden/extract_golden.py).
step-0 ``ll``  = llgrad(X_obs) + x_prior(X_obs) [+ cov_prior for xcov]  (gprfopt.py:396-405)
``trueX`` ll   = llgrad at the true X with X_obs-derived blocks          (gprfopt.py:505-509)
Values are printed with 2 decimals => tolerance 0.006 absolute (~1e-9 relative).
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "gprf_results_golden.json")))["runs"]
TOL = 0.006


def _cov_prior_at_init(lscale):
    # gprfopt.py:324-331 with c = log(C0) = log(lscale) (cov_scale cancels: xc = x/cov_scale)
    c = np.array([np.log(lscale)])
    r = (c + 1.0) / 10.0
    return -.5 * np.sum(r ** 2) - .5 * len(c) * np.log(2 * np.pi * 100.0)


def _cases(ntrains):
    return [pytest.param(r, id=r["dir"][:48] + "_" + r["task"] + str(r["init_seed"]))
            for r in GOLD if r["ntrain"] in ntrains and r["nblocks"] > 1]


@pytest.mark.parametrize("run", _cases((2000, 5000)))
def test_oracle_reproduces_golden(run, golden_data):
    sd = golden_data(run["ntrain"], run["nblocks"], run["local_dist"])
    xp_obs = sd.x_prior(sd.X_obs.flatten())[0]
    assert abs(xp_obs - run["step0_xprior"]) < 1e-6
    if run["init_seed"] == -9999:       # --init_true: starts from SX, blocks recomputed on SX
        gp = sd.build_gprf(local_dist=run["local_dist"])
        gp.update_X(sd.SX)
        ll0 = gp.llgrad()[0] + sd.x_prior(sd.SX.flatten())[0]
    else:
        ll0 = sd.build_gprf(local_dist=run["local_dist"]).llgrad()[0] + xp_obs
    if run["task"] == "xcov":
        ll0 += _cov_prior_at_init(sd.cov.dfn_params[0])
    assert abs(ll0 - run["step0_ll"]) < TOL
    if run.get("trueX_ll") is not None:
        llt = sd.build_gprf(X=sd.SX, local_dist=run["local_dist"]).llgrad()[0]
        assert abs(llt - run["trueX_ll"]) < TOL
        assert abs(sd.x_prior(sd.SX.flatten())[0] - run["trueX_xprior"]) < 1e-3


@pytest.mark.slow
@pytest.mark.parametrize("run", [pytest.param(r, id=r["dir"][:48]) for r in GOLD
                                 if r["ntrain"] == 10000 and r["nblocks"] == 100 and r["init_seed"] == -1])
def test_oracle_reproduces_golden_readme_config(run, golden_data):
    """BASELINE cfg 2: n=10000, 100 blocks, local GP and GPRF (342 edges)."""
    sd = golden_data(10000, 100, run["local_dist"])
    gp = sd.build_gprf(local_dist=run["local_dist"])
    assert len(gp.neighbors) == (342 if run["local_dist"] < 1.0 else 0)
    ll0 = gp.llgrad()[0] + sd.x_prior(sd.X_obs.flatten())[0]
    assert abs(ll0 - run["step0_ll"]) < TOL
    llt = sd.build_gprf(X=sd.SX, local_dist=run["local_dist"]).llgrad()[0]
    assert abs(llt - run["trueX_ll"]) < TOL
