import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from gprf_b200.synthetic import readme_dataset
sd = readme_dataset(ntrain=2000, nblocks=9)
rng = np.random.RandomState(1)
X1 = sd.X_obs.copy()
X2 = X1 + 1e-4 * rng.randn(*X1.shape)
X3 = X1 + 1e-9 * rng.randn(*X1.shape)
for fnt in (8, 0):
    g = sd.build_gprf(local_dist=0.1)
    g.set_fused_nt(fnt)
    res = []
    for X in (X1, X2, X1, X3, X1, X1):
        g.update_X(X)
        ll, gx, _ = g.llgrad(grad_X=True)
        res.append((ll, gx.copy()))
    base = res[0]
    for k in (2, 4, 5):
        print("fused_nt", fnt, "repeat", k, "ll equal", res[k][0] == base[0], "grad equal", np.array_equal(res[k][1], base[1]),
              "max grad diff", np.abs(res[k][1] - base[1]).max())
    g.close()
