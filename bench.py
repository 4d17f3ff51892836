#!/usr/bin/env python
"""Benchmark of the GPRF llgrad hot path (BASELINE.json metric: llgrad evals/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

One "step" = one objective-and-gradient evaluation (GPRF.llgrad with grad_X, as one
L-BFGS function evaluation of gprfopt.py:377-417) of the named workload.

  value      device-resident: X already in HBM, result left in HBM; CUDA events per step on
             the launching stream, L2 flushed between steps, max over ranks.
  e2e        the call a user makes: GPRF.update_X(X_host) + GPRF.llgrad(grad_X=True) returning
             host numpy arrays - host block assignment, H2D, kernels, (all-reduce), D2H.
  roofline   dominant kernel family, algorithmic FP64 flops / measured device time against
             the FP64 tensor (DMMA) peak measured live with a cuBLAS DGEMM.
  cpu_baseline   the CPU oracle (a port of the reference's algorithm) timed on this host.

N > 1 (torchrun, one rank per GPU): the SAME problem is sharded by units over the ranks
(strong scaling); partial [ll, grad] vectors are summed with one NCCL all-reduce.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DY = 50
POOL = 8                         # perturbed copies of X cycled through by the timed steps
FAMILY_FLOPS = {                 # algorithmic flops per unit of s points (sum = s^3 + 4 s^2 dy)
    "potrf": lambda s, dy: s ** 3 / 3.0 + s ** 2 * dy,      # dpotrf + forward half of dpotrs
    "trtri": lambda s, dy: s ** 3 / 3.0,                    # dpotri, triangular inverse
    "alpha": lambda s, dy: s ** 2 * dy,                             # back half of dpotrs
    "kinv_grad": lambda s, dy: s ** 3 / 3.0 + 2.0 * s ** 2 * dy,    # dpotri's U U^T + Alpha Alpha^T for G
}


# --------------------------------------------------------------------------- workloads
_README_SD = []


def make_workload(name):
    """Returns dict(X, Y, block_fn, block_idxs, neighbors, cov, noise_var, grad_cov, desc)."""
    from gprf_b200 import GPCov, Blocker, grid_centers
    from gprf_b200.synthetic import readme_dataset, SampledData
    if name in ("cfg2", "cfg3"):
        if not _README_SD:
            _README_SD.append(readme_dataset(ntrain=10000, nblocks=100, ntest=500, yd=DY, seed=0))
        sd = _README_SD[0]
        return dict(X=sd.X_obs, Y=sd.SY, block_fn=sd.reblock, block_idxs=sd.block_idxs, neighbors=sd.neighbors,
                    cov=sd.cov, noise_var=sd.noise_var, grad_cov=(name == "cfg3"), jiggle=0.003, clip01=False,
                    desc="gprfopt README config: n=10000 yd=50 nblocks=100 (342 edges) lscale=0.06 obs_std=0.02 "
                         "seed=0, X=X_obs, task=%s" % ("xcov" if name == "cfg3" else "x"))
    if name == "cfg1":
        sd = SampledData(noise_var=0.01, n=2500, ntrain=2000, lscale=0.06, obs_std=0.006, yd=DY, seed=0)
        sd.set_centers(grid_centers(20))
        return dict(X=sd.X_obs, Y=sd.SY, block_fn=sd.reblock, block_idxs=sd.block_idxs, neighbors=sd.neighbors,
                    cov=sd.cov, noise_var=0.01, grad_cov=False, jiggle=0.003,
                    desc="gprfopt synthetic n=2000 yd=50 lscale=0.06 nblocks=20 (25 blocks, 72 edges) task=x")
    if name == "cfg5":
        n = 200000
        rng = np.random.RandomState(0)
        X = rng.rand(n, 2)
        Y = rng.randn(n, DY)       # exact GP sampling at n=200k is infeasible; timing is data independent
        bl = Blocker(grid_centers(400))
        lscale = 6.0 / np.sqrt(n)
        return dict(X=X, Y=Y, block_fn=bl.block_clusters, block_idxs=bl.block_clusters(X), neighbors=bl.neighbors(),
                    cov=GPCov([1.0], [lscale, lscale], "euclidean", "se"), noise_var=0.01, grad_cov=False, jiggle=0.0005,
                    desc="synthetic n=200000 yd=50 400 blocks (~500 pts, 1482 edges) lscale=6/sqrt(n) Y=randn task=x")
    if name == "cfg4":
        # BASELINE configs[3]: sorted_isc.npy is absent from the reference mount, so a synthetic
        # catalogue of the same kind: events clustered along random "faults"
        # (cf. sample_crazy_lines, synthetic.py:35-50) in lon 60..100, lat 20..50, depth ~ Exp(30 km).
        from gprf_b200 import pdtree_cluster
        n, nf = 100000, 60
        rng = np.random.RandomState(0)
        a = np.column_stack([rng.uniform(62, 98, nf), rng.uniform(22, 48, nf)])
        d = rng.randn(nf, 2)
        d = d / np.linalg.norm(d, axis=1)[:, None] * rng.uniform(2.0, 6.0, nf)[:, None]
        f = rng.randint(0, nf, n)
        t = rng.rand(n)
        ll = a[f] + t[:, None] * d[f] + 0.3 * rng.randn(n, 2)
        X = np.column_stack([np.clip(ll[:, 0], 60, 100), np.clip(ll[:, 1], 20, 50),
                             np.clip(rng.exponential(30.0, n), 0, 700)])
        Y = rng.randn(n, DY)
        idxs, reblock = pdtree_cluster(X, blocksize=210)
        return dict(X=X, Y=Y, block_fn=reblock, block_idxs=idxs, neighbors=None, threshold=0.6,
                    cov=GPCov([1.0], [40.0, 40.0], "lld", "matern32"), noise_var=0.1, grad_cov=True, jiggle=0.01,
                    desc="seismic-style synthetic catalogue n=100000 (60 faults) lld+matern32 l=40km nv=0.1 yd=50 "
                         "pdtree blocksize=210 threshold=0.6 Y=randn task=xcov")
    raise SystemExit("unknown workload %r" % name)


def unit_sizes(block_idxs, neighbors):
    b = np.array([len(x) for x in block_idxs], dtype=np.float64)
    e = np.array(neighbors, dtype=np.int64).reshape(-1, 2)
    return np.concatenate([b, b[e[:, 0]] + b[e[:, 1]]]) if len(e) else b


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- reference / CPU arm
def oracle_gprf(wl, mode="fast"):
    from oracle.gprf_oracle import OracleGPRF
    from oracle.kernels import GPCov as OCov
    c = wl["cov"]
    return OracleGPRF(wl["X"], wl["Y"], None, OCov(c.wfn_params, c.dfn_params, c.dfn_str, c.wfn_str),
                      wl["noise_var"], block_idxs=wl["block_idxs"], neighbors=wl["neighbors"], mode=mode)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(wl, budget_s=20.0):
    """Oracle port on this host in the reference's own parallel mode (Pool(cpu_count) over units,
    gprf.py:218-233), 1 BLAS thread per worker - the fastest way the reference's algorithm runs on
    CPU (1 process x all BLAS threads is ~5x slower on these 100-200 point units).  Bounded sample:
    one warm-up evaluation, then whole evaluations until ~budget_s seconds are spent."""
    o = oracle_gprf(wl)
    kw = dict(grad_X=True, grad_cov=wl["grad_cov"])
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    try:
        t0 = time.perf_counter()
        o.llgrad(parallel=True, **kw)
        first = time.perf_counter() - t0
        ts = []
        while not ts or (sum(ts) + first + min(ts) < budget_s and len(ts) < 5):
            t0 = time.perf_counter()
            o.llgrad(parallel=True, **kw)
            ts.append(time.perf_counter() - t0)
    finally:
        if limiter is not None:
            limiter.restore_original_limits()
    sec = float(np.median(ts))
    modes = {}
    try:        # (i) one process, BLAS free to use every core (vectorised oracle)
        o.llgrad(**kw)
        t0 = time.perf_counter()
        o.llgrad(**kw)
        modes["one_process_all_blas_threads"] = {"sec_per_eval": time.perf_counter() - t0, "blas_threads": blas_threads(),
                                                 "sample": "1 full eval after 1 warm-up"}
    except Exception as exc:        # noqa: BLE001
        modes["one_process_all_blas_threads"] = {"error": str(exc)}
    try:        # (ii) the reference's loop structure (one kernel_deriv row call per point and dimension,
        #      gprf.py:553-561): a bounded sample of units, scaled to the whole evaluation by flops
        of = oracle_gprf(wl, mode="faithful")
        sizes = unit_sizes(wl["block_idxs"], wl["neighbors"])
        w = sizes ** 3 + 4 * sizes ** 2 * DY
        nb = of.n_blocks
        pick_b = list(range(0, nb, max(1, nb // 6)))[:6]
        pick_e = list(range(0, len(of.neighbors), max(1, len(of.neighbors) // 6)))[:6]
        t0 = time.perf_counter()
        for b_ in pick_b:
            of.llgrad_unary(b_, **kw)
        for e_ in pick_e:
            of.llgrad_joint(*of.neighbors[e_], **kw)
        dt = time.perf_counter() - t0
        wsum = w[pick_b].sum() + w[[nb + e_ for e_ in pick_e]].sum()
        modes["faithful_row_loops_serial"] = {"sec_per_eval": dt * w.sum() / wsum,
                                              "sample": "%d blocks + %d pairs, serial, scaled by flops (%.1f%% of the eval)"
                                                        % (len(pick_b), len(pick_e), 100 * wsum / w.sum())}
    except Exception as exc:        # noqa: BLE001
        modes["faithful_row_loops_serial"] = {"error": str(exc)}
    return {"value": 1.0 / sec, "unit": "evals/s", "cores": cores, "kind": "port", "modes": modes,
            "reference_logged": {"sec_per_eval": 7.31, "hardware": "unstated CPU, serial (--parallel off)",
                                 "source": "gprf_results.tgz: 10000_10500_100_0.060000_0.020000_0.1000_50_l-bfgs-b_x_-1_"
                                           "0.0100_s0_gprf0/results.txt (README configuration)"},
            "sample": "%d full evals after 1 warm-up (median), all %d units, grad_X" % (len(ts), o.n_blocks + len(o.neighbors)),
            "sec_per_eval": sec,
            "note": "oracle port, Pool(%d) x 1 BLAS thread (the reference's --parallel mode)" % cores}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores.  The reference itself
    (Python 2 + un-vendored treegp) cannot run, so this is the oracle port, using the reference's
    own parallel mode (Pool(cpu_count) over units, gprf.py:218-233) with 1 BLAS thread per worker."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload)
    o = oracle_gprf(wl)
    kw = dict(grad_X=True, grad_cov=wl["grad_cov"])
    cores = os.cpu_count() or 1
    sizes = unit_sizes(wl["block_idxs"], wl["neighbors"])
    w = sizes ** 3 + 4 * sizes ** 2 * DY
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    # bounded sample: keep every step under ~20 s by evaluating a prefix of the edge list
    t0 = time.perf_counter()
    o.llgrad_joint(*o.neighbors[0], **kw)
    est = (time.perf_counter() - t0) * w.sum() / w[o.n_blocks] / max(1, cores)
    frac = 1.0
    if est > 20.0:
        keep = max(cores, int(len(o.neighbors) * 20.0 / est))
        used = np.zeros(len(sizes), dtype=bool)
        used[:o.n_blocks] = True
        used[o.n_blocks:o.n_blocks + keep] = True
        frac = w[used].sum() / w.sum()
        o.neighbors = o.neighbors[:keep]
        o.compute_neighbor_count()
    for _ in range(max(1, args.warmup if frac == 1.0 else 1)):
        o.llgrad(parallel=True, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.llgrad(parallel=True, **kw)
    sec = (time.perf_counter() - t0) / args.steps / frac
    sample = ("full eval per step" if frac == 1.0 else
              "%.1f%% of the eval's flops per step (all blocks + first %d edges), scaled by flops"
              % (100 * frac, len(o.neighbors)))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": "GPRF llgrad evals/sec", "value": val, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": wl["desc"]},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "reference is Python 2 + un-vendored treegp (cannot run); oracle port in the "
                                     "reference's --parallel mode: Pool(%d) x 1 BLAS thread" % cores},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def dgemm_peak_tflops(torch, n=6144, reps=4):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / best * 1e-9


class Runner(object):
    """One workload on this rank's GPU: device-resident steps, end-to-end steps, profiling."""

    def __init__(self, torch, dist, wl, rank, world, local_rank):
        from gprf_b200 import GPRF
        from gprf_b200 import _lib
        self.torch, self.dist, self.wl, self.rank, self.world = torch, dist, wl, rank, world
        self.n, self.dx = wl["X"].shape
        self.dev = torch.device("cuda", local_rank)
        self.g = GPRF(wl["X"], wl["Y"], wl["block_fn"], wl["cov"], wl["noise_var"], block_idxs=wl["block_idxs"],
                      neighbors=wl["neighbors"], device=local_rank, neighbor_threshold=wl.get("threshold", 1e-3),
                      unit_shard=(rank, world) if world > 1 else None)
        if wl["neighbors"] is None:                      # threshold edges (gprf.py:119-150), computed on the device
            wl["neighbors"] = list(self.g.neighbors)
            wl["desc"] += " (%d blocks, %d edges)" % (self.g.n_blocks, len(wl["neighbors"]))
        self.outlen = 1 + _lib.MAX_NCOV + self.n * self.dx
        # Every step evaluates a DIFFERENT X (as consecutive L-BFGS iterates do): a pool of perturbed
        # copies, so that points change block, block sizes change and nothing structural can be
        # carried over from the previous step.
        rng = np.random.RandomState(1234)
        scale = wl.get("jiggle", 0.0)
        self.pool_h = [np.ascontiguousarray(wl["X"], dtype=np.float64)]
        for _ in range(POOL - 1 if scale > 0 else 0):
            Xp = np.array(wl["X"], dtype=np.float64)
            Xp[:, :2] += scale * rng.randn(self.n, 2)
            if wl.get("clip01"):
                Xp = np.clip(Xp, 0.0, 1.0)
            self.pool_h.append(np.ascontiguousarray(Xp))
        self.pool_d = [torch.tensor(x, dtype=torch.float64, device=self.dev) for x in self.pool_h]
        self.step_no = 0
        self.Xd = self.pool_d[0]
        # [status | ll, grad theta (5), gradX]: at N > 1 the evaluation's status word travels with the
        # results through the all-reduce (no host round trip before the collective)
        self.out_s = torch.zeros(self.outlen + 1, dtype=torch.float64, device=self.dev)
        self.out = self.out_s[1:]
        self.status_h = torch.zeros(1, dtype=torch.float64).pin_memory()
        self.Xe2e = torch.empty((self.n, self.dx), dtype=torch.float64, device=self.dev)
        self.sync_redos = 0
        self.Xh = torch.empty((self.n, self.dx), dtype=torch.float64).pin_memory()
        self.outh = torch.empty(self.outlen, dtype=torch.float64).pin_memory()
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.grad_cov = wl["grad_cov"]
        self.reblock = self.g._device_part is not None     # a step = update_X (re-blocking) + llgrad

    def flush_l2(self):
        self.flush_buf.zero_()

    def device_step(self, reblock=False, Xd=None, force_sync=False):
        """One evaluation with X resident in HBM.  N = 1: gprf_llgrad_device (one host round trip for the
        status word).  N > 1: gprf_llgrad_device_nosync + ONE all-reduce of [status | results]; the caller
        synchronises and calls check_status()."""
        st = self.torch.cuda.current_stream(self.dev)
        self.step_no += 1
        if Xd is None:
            Xd = self.pool_d[self.step_no % len(self.pool_d)] if reblock else self.pool_d[0]
        self.Xd = Xd
        if self.world == 1 or force_sync:
            self.g.llgrad_device(Xd.data_ptr(), self.out.data_ptr(), st.cuda_stream,
                                 grad_X=True, grad_cov=self.grad_cov, reblock=reblock)
            if self.world > 1:
                self.out_s[0] = 0.0
                self.dist.all_reduce(self.out_s)
                self.status_h.zero_()
            return
        self.g.llgrad_device(Xd.data_ptr(), self.out.data_ptr(), st.cuda_stream, grad_X=True, grad_cov=self.grad_cov,
                             reblock=reblock, status_dev_ptr=self.out_s.data_ptr())
        self.dist.all_reduce(self.out_s)
        self.status_h.copy_(self.out_s[:1], non_blocking=True)

    def check_status(self):
        """After the stream has drained (N > 1): a non-zero reduced status means some rank's evaluation has
        to be redone by the synchronous path (jitter rule / tile pipeline) - on every rank."""
        if self.world > 1 and self.status_h[0].item() != 0.0:
            st = self.torch.cuda.current_stream(self.dev)
            self.g.llgrad_device(self.Xd.data_ptr(), self.out.data_ptr(), st.cuda_stream,
                                 grad_X=True, grad_cov=self.grad_cov, reblock=False)
            self.out_s[0] = 0.0
            self.dist.all_reduce(self.out_s)
            st.synchronize()
            self.sync_redos += 1

    def timed_device_steps(self, k):
        torch = self.torch
        total = 0.0
        launches = 0
        for _ in range(k):
            self.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.device_step(reblock=self.reblock)
            e1.record()
            e1.synchronize()
            self.check_status()
            total += e0.elapsed_time(e1)
            launches += self.g.last_timing()[1] + (1 if self.world > 1 else 0)
        return total, launches

    def e2e_step(self, X_host=None):
        """update_X + llgrad through host buffers (the call a user makes), a different X every step."""
        torch = self.torch
        if X_host is None:
            self.step_no += 1
            X_host = self.pool_h[self.step_no % len(self.pool_h)] if self.g.block_fn is not None else self.pool_h[0]
        if self.world == 1:
            self.g.update_X(X_host)
            return self.g.llgrad(grad_X=True, grad_cov=self.grad_cov)
        self.g.update_X(X_host)
        self.Xh.numpy()[...] = X_host
        self.Xe2e.copy_(self.Xh, non_blocking=True)
        self.device_step(reblock=self.g._device_part is not None, Xd=self.Xe2e)
        self.outh.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        self.check_status()
        return self.outh[0].item()

    def e2e_bytes(self):
        """Bytes that cross PCIe in one e2e step.  With the on-device partitioner: X down; the block
        bounds (to size the units), the Cholesky status word and [ll, grad theta, gradX] up.  Unit
        descriptors go down again only when some block changed size (not the case for a repeated X).
        Without it the host-computed block lists and descriptors are uploaded every step."""
        nb, ne = len(self.wl["block_idxs"]), len(self.wl["neighbors"])
        h2d = self.n * self.dx * 8
        d2h = (1 + 5 + self.n * self.dx) * 8 + 4
        ev, fb, _ = self.g.resident_stats()
        if ev > 0 and fb == 0:
            pass                                      # resident path: X down, [ll, grad] + status word up, nothing else
        elif self.g._device_part is not None:
            d2h += (nb + 1) * 8
        else:
            h2d += self.n * 8 + self.n * 4 + (nb + 1) * 16 + (nb + ne) * (176 + 4)
        return h2d, d2h

    def family_profile(self, reps=3):
        self.g.set_profiling(True)
        acc = {}
        for _ in range(reps):
            self.flush_l2()
            self.device_step(reblock=self.reblock, force_sync=True)     # the per-family events are resolved by the synchronous entry
            for k, (ms, nl) in self.g.family_timing().items():
                a = acc.setdefault(k, [0.0, 0])
                a[0] += ms / reps
                a[1] = nl
        self.g.set_profiling(False)
        return acc


FUSED_NT = int(os.environ.get("GPRF_FUSED_NT", "0"))   # library default (include/gprf_b200.h): tile pipeline only
FUSED_MIXED_NT = int(os.environ.get("GPRF_FUSED_MIXED_NT", str(min(FUSED_NT, 4))))


def lbfgs_full_run():
    """BASELINE.json configs[1] end to end: the whole L-BFGS optimisation of the README configuration
    through gprf_b200/gprfopt.py (py3 mirror of the reference driver, same scipy options), compared with
    the run the reference logged for the same seed (tests/golden/gprf_trajectories_golden.json)."""
    import tempfile
    from gprf_b200 import gprfopt
    from gprf_b200.synthetic import readme_dataset
    name = "10000_10500_100_0.060000_0.020000_0.1000_50_l-bfgs-b_x_-1_0.0100_s0_gprf0"
    gold = None
    gpath = os.path.join(ROOT, "tests", "golden", "gprf_trajectories_golden.json")
    if os.path.exists(gpath):
        gold = [r for r in json.load(open(gpath))["runs"] if r["dir"] == name][0]["steps"]
    sd = readme_dataset(ntrain=10000, nblocks=100, ntest=500, yd=DY, seed=0)
    gp = sd.build_gprf(local_dist=0.1)
    for _ in range(3):
        gp.llgrad(grad_X=True)                       # warm the context outside the timed run
    walls = []
    for _ in range(2):          # the same deterministic run twice: host jitter (scipy's L-BFGS core is ~half of it)
        gp.update_X(sd.X_obs)
        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            log = gprfopt.do_optimization(d, gp, sd.X_obs, None, sd, save_steps=False)
            walls.append(time.perf_counter() - t0)
    wall = min(walls)
    out = {"config": name, "evals": len(log), "wall_s": wall, "wall_s_runs": walls, "evals_per_s": len(log) / wall,
           "final_objective": log[-1][2], "best_objective": max(l[2] for l in log),
           "note": "scipy L-BFGS-B ftol 1e-6 maxiter 200 (gprfopt.py:418); host numpy in/out every evaluation, "
                   "step_*.npy dumps off"}
    if gold:
        nsame = 0
        for (st, _t, ll), g in zip(log, gold):
            if abs(ll - g[1]) > 0.011:
                break
            nsame += 1
        out["reference_log"] = {"evals": len(gold), "final_objective": gold[-1][1],
                                "seconds_logged_by_reference": 650.03,
                                "leading_evals_identical_to_2_decimals": nsame}
    gp.close()
    return out


RESIDENT_SOURCES = ("resident.cuh", "covfn.cuh", "tile_gemm.cuh", "smem_chol.cuh")


def kernel_sources_sha16(files=None):
    """Hash of device code - the stamp of profiles/ncu_traffic.json.  Every kernel lives in a .cuh header (the
    .cu files hold the host plan and the C-ABI); ``files`` = the headers one kernel is compiled from (k_resident:
    RESIDENT_SOURCES, the include closure of gprf_resident.cu), default all of them."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "gprf_b200", "csrc")
    for fn in sorted(os.listdir(csrc)):
        if fn.endswith(".cuh") and (files is None or fn in files):
            with open(os.path.join(csrc, fn), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def roofline_from_profile(fam, sizes_local, dy, peak_tflops, peak_note):
    """Dominant kernel family of one evaluation.  Units of up to FUSED_NT tiles run the whole pipeline
    in k_unit_fused (algorithmic flops s^3 + 4 s^2 dy each); larger ones go through the per-family
    tile kernels, whose algorithmic flops are split as in FAMILY_FLOPS."""
    nts = np.ceil(sizes_local / 64.0)
    # library rule (include/gprf_b200.h, gprf_set_fused_nt): once some unit needs the tile pipeline,
    # only units of up to min(FUSED_NT, 4) tiles stay in the fused kernel
    thr = min(FUSED_NT, FUSED_MIXED_NT) if nts.size and nts.max() > FUSED_NT else FUSED_NT
    fused = nts <= thr
    big = sizes_local[~fused]
    merged = {"potrf": [fam["potrf_diag"][0] + fam["potrf_panel"][0], fam["potrf_diag"][1] + fam["potrf_panel"][1]],
              "trtri": fam["trtri"], "alpha": fam["alpha"], "kinv_grad": fam["kinv_grad"],
              "unit_fused": fam["unit_fused"],
              # resident path: ONE kernel (k_resident) evaluates every unit out of shared memory
              "resident": [fam["res_blocks"][0] + fam["res_pairs"][0], fam["res_blocks"][1] + fam["res_pairs"][1]]}
    name = max(merged, key=lambda k: merged[k][0])
    ms, nl = merged[name]
    if name == "resident":
        flops = float(np.sum(sizes_local ** 3 + 4 * sizes_local ** 2 * dy))
    elif name == "unit_fused":
        sf = sizes_local[fused]
        flops = float(np.sum(sf ** 3 + 4 * sf ** 2 * dy))
    else:
        flops = float(np.sum(FAMILY_FLOPS[name](big, dy)))
    achieved = flops / (ms * 1e-3) * 1e-12 if ms > 0 else 0.0
    # every tile-pipeline family against the same peak (reference-algorithm flops, also with factor reuse on)
    fam_frac = {}
    for k in FAMILY_FLOPS:
        if merged[k][0] > 0 and big.size:
            fam_frac[k] = round(float(np.sum(FAMILY_FLOPS[k](big, dy))) / (merged[k][0] * 1e-3) * 1e-12 / peak_tflops, 4)
    return {"bound": "tensor", "families_frac_of_peak": fam_frac, "kernel": name, "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / peak_tflops, "traffic": None, "peak_source": peak_note,
            "launches_per_eval": nl, "ms_per_eval": ms,
            "families_ms": dict((k, round(v[0], 4)) for k, v in fam.items())}


def measure(torch, dist, args, wl_name, rank, world, local_rank, steps, warmup, with_cpu):
    wl = make_workload(wl_name)
    R = Runner(torch, dist, wl, rank, world, local_rank)
    sizes = unit_sizes(wl["block_idxs"], wl["neighbors"])
    flops_eval = float(np.sum(sizes ** 3 + 4 * sizes ** 2 * DY))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a few 100 ms to start: cover warm-up too
    # Every loop below runs the SAME number of steps on every rank (each step holds an all-reduce):
    # counts are fixed or derived from all-reduced quantities, never from a rank's own clock.
    t_w = time.perf_counter()
    for _ in range(max(3, warmup)):
        R.device_step(reblock=R.reblock)
    torch.cuda.synchronize()
    tw = torch.tensor([(time.perf_counter() - t_w) / max(3, warmup)], dtype=torch.float64, device=R.dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    n_extra = int(min(400, max(0, 0.8 / max(tw.item(), 1e-6))))      # ~0.8 s more of untimed warm-up
    for _ in range(n_extra):
        R.device_step(reblock=R.reblock)
    barrier()
    total_ms, launches = R.timed_device_steps(steps)
    barrier()
    t = torch.tensor([total_ms], dtype=torch.float64, device=R.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / steps
    # keep the load on (untimed) until the clock sampler has seen it: ~1.5 s worth of steps
    for _ in range(int(min(2000, max(3, 1500.0 / max(ms_per_step, 1e-3))))):
        R.device_step(reblock=R.reblock)
    torch.cuda.synchronize()
    clocks = sampler.stop()

    # end to end through the public host API
    for _ in range(2):
        R.e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        R.e2e_step()
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=R.dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_sec = te.item()
    h2d, d2h = R.e2e_bytes()

    peak = dgemm_peak_tflops(torch)
    peak_note = ("measured live: cuBLAS DGEMM 6144^3 fp64 on this GPU (MEASURED_PEAKS.json has no fp64 entry; raw DMMA "
                 "issue peak measured at 37.1 TFLOP/s, profiles/r01_fp64_peak_dmma_dfma.txt)")
    fam = R.family_profile()
    if world > 1:
        from gprf_b200.dist import shard_units
        from gprf_b200.gprf import _blocks_to_csr
        ptr, _ = _blocks_to_csr(wl["block_idxs"])
        nominal = R.g.resident_stats()[0] > 0          # small-block structures are split on nominal sizes
        mask = shard_units(ptr, np.asarray(wl["neighbors"]), rank, world, nominal=nominal).astype(bool)
        sizes_local = sizes[mask]
    else:
        sizes_local = sizes
    roof = roofline_from_profile(fam, sizes_local, DY, peak, peak_note)
    # DRAM bytes of the dominant kernel, from an ncu --set full capture of THIS source state (1 GPU):
    # profiles/ncu_traffic.json is stamped with a hash of the kernel sources and ignored on a mismatch.
    # On the resident path the dominant kernel's one launch IS the evaluation (the other three launches
    # move < 1 MB), so the figure is per evaluation.
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        ent = tj.get(wl_name, {}).get(roof["kernel"])
        if ent and world == 1:
            sha = kernel_sources_sha16(RESIDENT_SOURCES if roof["kernel"] == "resident" else None)
            if ent.get("sources_sha16", sha) == sha:
                roof["traffic"] = ent["bytes"]
                roof["traffic_unit"] = ent.get("unit", "bytes per launch (dram read + write)")
                roof["traffic_source"] = ent["source"]
                if "algorithmic_bytes" in ent:
                    roof["traffic_algorithmic_bytes"] = ent["algorithmic_bytes"]
            else:
                roof["traffic"] = None
                roof["traffic_source"] = ("profiles/ncu_traffic.json was captured from other kernel sources (%s); "
                                          "re-run scripts/gpu_job_ncu.sh" % ent.get("sources_sha16"))
    except (IOError, OSError, ValueError):
        pass
    roof["eval_tflops"] = flops_eval / (ms_per_step * 1e-3) * 1e-12
    roof["eval_frac_of_peak_all_gpus"] = roof["eval_tflops"] / (peak * world)
    ar_ms = None
    if world > 1:                        # the one collective of the path, timed alone (CUDA events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(10):
            dist.all_reduce(R.out_s)
        e1.record()
        e1.synchronize()
        ta = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=R.dev)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        ar_ms = ta.item()
    roof["factor_reuse"] = dict(zip(("pair_units", "tile_tasks_saved"), R.g.factor_reuse_stats()))
    if ar_ms is not None:
        roof["allreduce_ms"] = ar_ms
        roof["allreduce_bytes"] = int(R.out_s.numel() * 8)
        roof["sync_redos"] = R.sync_redos
    ev, fb, _ = R.g.resident_stats()
    roof["resident_path"] = {"evaluations": ev, "handed_to_tile_pipeline": fb}
    res = {"wl": wl, "reblock": R.reblock, "ms_per_step": ms_per_step, "value": 1e3 / ms_per_step, "launches": launches // steps,
           "clocks": clocks, "e2e": {"value": 1.0 / e2e_sec, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                                     "d2h_bytes_per_step": d2h, "ms_per_step": e2e_sec * 1e3},
           "roofline": roof, "flops_per_eval": flops_eval}
    if with_cpu and rank == 0 and world == 1:
        try:
            res["cpu_baseline"] = cpu_baseline(wl)
        except Exception as exc:        # noqa: BLE001  (a broken host pool must not cost the GPU line)
            res["cpu_baseline"] = {"value": None, "unit": "evals/s", "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": "failed: %s: %s" % (type(exc).__name__, exc)}
    R.g.close()
    del R
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--no-n200k", action="store_true", help="skip the extra n=200k measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg1 / cfg3 / cfg4 sub-lines")
    ap.add_argument("--no-lbfgs", action="store_true", help="skip the full L-BFGS run of the README configuration")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version
    # banner, nvcc, library chatter) is sent to stderr for the rest of the run
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"           # keep stdout to the one JSON line
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=150))   # a hang must not eat the GPU budget
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()

    r = measure(torch, dist, args, args.workload, rank, world, local_rank, args.steps, args.warmup,
                with_cpu=not args.no_cpu)
    line = {"metric": "GPRF llgrad evals/sec", "value": r["value"], "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": r["wl"]["desc"], "parallelism": "units sharded over %d GPU(s), 1 all-reduce" % world,
                       "l2": "flushed between timed steps (256 MiB write)",
                       "step": "re-blocking of X on the device + llgrad(grad_X)" if r.get("reblock") else "llgrad(grad_X)",
                       "flops_per_eval": r["flops_per_eval"]},
            "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["launches"] * args.steps,
            "roofline": r["roofline"]}
    if "cpu_baseline" in r:
        line["cpu_baseline"] = r["cpu_baseline"]
    # The two extra measurements below must never cost the headline line (a failure on ONE rank of
    # a multi-rank n=200k run would still hang the others in the collective: NCCL's timeout ends it).
    if world == 1 and args.workload == "cfg2" and not args.no_lbfgs:
        try:
            line["lbfgs_full_run"] = lbfgs_full_run()
        except Exception as exc:        # noqa: BLE001
            line["lbfgs_full_run"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if world == 1 and args.workload == "cfg2" and not args.no_extra:
        # the other BASELINE.json configurations, one short measurement each (parity-tested in tests/)
        line["configs"] = {}
        for name in ("cfg1", "cfg3", "cfg4"):
            try:
                kx = max(3, min(args.steps, 10))
                rx = measure(torch, dist, args, name, rank, world, local_rank, kx, 3, with_cpu=False)
                line["configs"][name] = {"workload": rx["wl"]["desc"], "value": rx["value"], "unit": "evals/s",
                                         "ms_per_step": rx["ms_per_step"], "steps": kx, "e2e": rx["e2e"],
                                         "roofline": rx["roofline"], "clocks": rx["clocks"],
                                         "gpu_launches_per_step": rx["launches"]}
                if name == "cfg4":
                    line["configs"][name]["parity"] = ("lld + Matern-3/2: CUDA == oracle is tested, the oracle itself is "
                                                       "parity UNPINNED (treegp source and sorted_isc.npy absent)")
            except Exception as exc:        # noqa: BLE001
                line["configs"][name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if not args.no_n200k and args.workload != "cfg5":
        try:
            k5 = max(2, min(args.steps, 5))
            r5 = measure(torch, dist, args, "cfg5", rank, world, local_rank, k5, 1, with_cpu=False)
            line["n200k"] = {"workload": r5["wl"]["desc"], "value": r5["value"], "unit": "evals/s",
                             "ms_per_step": r5["ms_per_step"], "steps": k5, "e2e": r5["e2e"],
                             "roofline": r5["roofline"], "clocks": r5["clocks"], "scaling": "strong"}
        except Exception as exc:        # noqa: BLE001
            line["n200k"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
