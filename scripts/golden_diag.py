"""Diagnostic: every dense-branch golden run, computed on the device vs logged."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_oracle_golden import _cov_prior_at_init
from gprf_b200 import grid_centers
from gprf_b200.synthetic import SampledData
GOLD = json.load(open("tests/golden/gprf_results_golden.json"))["runs"]
for ntrain in sorted(set(r["ntrain"] for r in GOLD if r["ntrain"] < 40000)):
    if len(sys.argv) > 1 and ntrain not in [int(a) for a in sys.argv[1:]]:
        continue
    sd = SampledData(noise_var=0.01, n=ntrain + 500, ntrain=ntrain, lscale=6.0 / np.sqrt(ntrain),
                     obs_std=2.0 / np.sqrt(ntrain), yd=50, seed=0, device=0)
    xp_obs = sd.x_prior(sd.X_obs.flatten())[0]
    for run in [r for r in GOLD if r["ntrain"] == ntrain]:
        sd.set_centers(grid_centers(run["nblocks"]))
        gp = sd.build_gprf(local_dist=run["local_dist"])
        if run["init_seed"] == -9999:
            gp.update_X(sd.SX)
            ll0 = gp.llgrad()[0] + sd.x_prior(sd.SX.flatten())[0]
        else:
            ll0 = gp.llgrad()[0] + xp_obs
        ne = len(gp.neighbors)
        gp.close()
        if run["task"] == "xcov":
            ll0 += _cov_prior_at_init(sd.cov.dfn_params[0])
        llt = None
        if run.get("trueX_ll") is not None:
            gt = sd.build_gprf(X=sd.SX, local_dist=run["local_dist"])
            llt = gt.llgrad()[0]
            gt.close()
        print(run["dir"][:60], "edges", ne, "step0 %.2f vs %.2f (%.3g)" % (ll0, run["step0_ll"], ll0 - run["step0_ll"]),
              "" if llt is None else "trueX %.2f vs %.2f (%.3g)" % (llt, run["trueX_ll"], llt - run["trueX_ll"]), "evals", run["n_evals"])
