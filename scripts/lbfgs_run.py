"""Full L-BFGS run of one of the reference's golden experiments through the CUDA-backed GPRF
(BASELINE.json configs[1]: "gprfopt.py README config ... full L-BFGS run on 1 B200"), compared
evaluation by evaluation with the trajectory the reference itself logged (results.txt inside
gprf_results.tgz -> tests/golden/gprf_trajectories_golden.json).

    python scripts/lbfgs_run.py [dir-prefix] [max_evals]
"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gprf_b200 import gprfopt
from gprf_b200.synthetic import readme_dataset

T = json.load(open(os.path.join(ROOT, "tests", "golden", "gprf_trajectories_golden.json")))["runs"]
which = sys.argv[1] if len(sys.argv) > 1 else "10000_10500_100_0.060000_0.020000_0.1000_50_l-bfgs-b_x_-1"
nev = int(sys.argv[2]) if len(sys.argv) > 2 else None
run = [r for r in T if r["dir"].startswith(which)][0]
sd = readme_dataset(ntrain=run["ntrain"], nblocks=run["nblocks"])
d = tempfile.mkdtemp()
t0 = time.time()
gp, log, rows = gprfopt.do_run(d, sd, local_dist=run["local_dist"], task=run["task"], init_seed=run["init_seed"],
                               max_evals=nev)
wall = time.time() - t0
gold = run["steps"]
print("# %s: %d evaluations (reference logged %d), optimisation %.2f s wall (log: %.2f s at the last step)"
      % (run["dir"], len(log), len(gold), wall, log[-1][1]))
worst = 0.0
for (st, sec, ll, c1, mad, xp), g in zip(rows, gold):
    rel = abs(ll - g[1]) / max(abs(g[1]), 1.0)
    worst = max(worst, rel)
    if st < 12 or st % 10 == 0 or st >= len(rows) - 3:
        print("%3d t=%7.2f ll %.2f ref %.2f rel %.1e | mad %.8f ref %.8f | xprior %.6f ref %.6f | lscale %.6f ref %.6f"
              % (st, sec, ll, g[1], rel, mad, g[3], xp, g[4], c1, g[2]))
print("# worst relative deviation of the objective over %d common evaluations: %.2e; final ll %.2f (reference %.2f)"
      % (min(len(rows), len(gold)), worst, rows[-1][2], gold[-1][1]))
