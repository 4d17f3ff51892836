"""Extract known-answer vectors from the reference's shipped result logs.

Run in the build container only (``/root/reference`` does not exist on the GPU
box); the output ``gprf_results_golden.json`` is committed.

Source: ``/root/reference/gprf_results.tgz`` - one ``results.txt`` per run,
written by gprfopt.py:486-488,513-515.  Columns (gprfopt_analyze.py:20-22):
``step time ll lscale_err mad xprior ...`` and a final ``trueX`` line.  Only
``*_gprf0`` runs (the GPRF objective, not the GPy baselines) are kept.

Directory name (gprfopt.py:588-597):
ntrain_n_nblocks_lscale_obsstd_localdist_yd_method_task_initseed_noisevar_sSEED_gprf0
"""
import json
import os
import sys
import tarfile

SRC = "/root/reference/gprf_results.tgz"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gprf_results_golden.json")
OUT_TRAJ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gprf_trajectories_golden.json")


def main():
    runs = []
    trajs = []
    with tarfile.open(SRC) as tf:
        for m in tf.getmembers():
            if not m.name.endswith("results.txt"):
                continue
            d = os.path.basename(os.path.dirname(m.name))
            if not d.endswith("_gprf0"):
                continue
            f = d.split("_")
            lines = tf.extractfile(m).read().decode().strip().split("\n")
            steps = [l.split() for l in lines if not l.startswith("trueX")]
            true = [l.split() for l in lines if l.startswith("trueX")]
            if not steps or steps[0][0] != "0":
                continue
            rec = {
                "dir": d,
                "ntrain": int(f[0]), "n": int(f[1]), "nblocks": int(f[2]),
                "lscale": float(f[3]), "obs_std": float(f[4]), "local_dist": float(f[5]),
                "yd": int(f[6]), "task": f[8], "init_seed": int(f[9]),
                "noise_var": float(f[10]), "seed": int(f[11][1:]),
                "step0_ll": float(steps[0][2]), "step0_xprior": float(steps[0][5]),
                "n_evals": len(steps),
                "t_first": float(steps[0][1]), "t_last": float(steps[-1][1]),
            }
            if true:
                rec["trueX_ll"] = float(true[0][2]) if true[0][2] != "-inf" else None
                rec["trueX_xprior"] = float(true[0][5])
            runs.append(rec)
            # full optimiser trajectory of the small runs and of the README configuration:
            # (step, ll, mean distance to the true X, x_prior) per objective evaluation
            if f[0] in ("2000", "5000") or (f[0] == "10000" and f[2] == "100"):
                trajs.append({"dir": d, "ntrain": rec["ntrain"], "nblocks": rec["nblocks"],
                              "local_dist": rec["local_dist"], "task": rec["task"], "init_seed": rec["init_seed"],
                              "seed": rec["seed"],
                              "steps": [[int(x[0]), float(x[2]), float(x[3]), float(x[4]), float(x[5])]
                                        for x in steps]})
    runs.sort(key=lambda r: (r["ntrain"], r["nblocks"], r["local_dist"], r["task"], r["init_seed"]))
    with open(OUT_TRAJ, "w") as fh:
        json.dump({"source": "gprf_results.tgz (davmre/gprf): results.txt rows, columns step ll lscale_ratio mad xprior",
                   "runs": trajs}, fh)
    print("wrote %d trajectories to %s" % (len(trajs), OUT_TRAJ))
    with open(OUT, "w") as fh:
        json.dump({"source": "gprf_results.tgz (davmre/gprf)", "runs": runs}, fh, indent=1)
    print("wrote %d runs to %s" % (len(runs), OUT))


if __name__ == "__main__":
    sys.exit(main())
