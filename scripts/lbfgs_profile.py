import cProfile, pstats, sys, os, time, tempfile
sys.path.insert(0, os.getcwd())
import bench
from gprf_b200 import gprfopt
from gprf_b200.synthetic import readme_dataset
sd = readme_dataset(ntrain=10000, nblocks=100, ntest=500, yd=50, seed=0)
gp = sd.build_gprf(local_dist=0.1)
for _ in range(3): gp.llgrad(grad_X=True)
t0=time.perf_counter()
for _ in range(20):
    gp.update_X(sd.X_obs); gp.llgrad(grad_X=True)
print("plain update_X+llgrad: %.3f ms" % ((time.perf_counter()-t0)/20*1e3))
with tempfile.TemporaryDirectory() as d:
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter()
    log = gprfopt.do_optimization(d, gp, sd.X_obs, None, sd, save_steps=False)
    wall = time.perf_counter() - t0
    pr.disable()
print(len(log), wall)
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
