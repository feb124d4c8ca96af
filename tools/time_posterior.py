#!/usr/bin/env python3
"""Likelihood throughput on POSTERIOR walkers (the committed equilibrated HD155358 ensemble, tiled to fill the GPU) -- the
workload the samplers' ESS/s is measured on -- with and without the cost-ordered item schedule (model option cost_order).
Usage: python tools/time_posterior.py [copies of the 28416-walker ensemble, default 4]"""
import json
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvtest as T
from rvel_mcmc_b200 import _abi

ctx = _abi.Context(0)
obs = T.load_vels("HD155358.vels")
oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
ens = np.load(os.path.join(ROOT, "tests", "golden", "hd155358_equilibrated_ensemble.npy"))
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rng = np.random.RandomState(3)
th = np.tile(ens, (copies, 1))
th = th[rng.permutation(len(th))]                    # walkers arrive in no particular order
W = len(th)
theta = torch.from_numpy(np.ascontiguousarray(th)).cuda()
logp = torch.empty(W, dtype=torch.float64, device="cuda"); st = torch.empty(W, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
ref = None
for opts in ({"cost_order": 0}, {"cost_order": 1}, {"cost_order": 0, "dense_output": 1}, {"cost_order": 1, "dense_output": 1}):
    for k in ("cost_order", "dense_output"):
        m.set_option(k, opts.get(k, 0))
    m.loglik_dev(oh, theta.data_ptr(), W, logp.data_ptr(), st.data_ptr(), s); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.loglik_dev(oh, theta.data_ptr(), W, logp.data_ptr(), st.data_ptr(), s); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    lp = logp.cpu().numpy(); sv = st.cpu().numpy()
    if ref is None:
        ref = (lp.copy(), sv.copy())
    same = bool(np.array_equal(sv, ref[1]) and (np.array_equal(lp, ref[0]) if not opts.get("dense_output") else np.abs(lp - ref[0])[sv == 0].max() < 1e-9))
    print(json.dumps({"walkers": W, "options": opts, "ms": best, "evals_per_s": W / best * 1e3, "ok_fraction": float((sv == 0).mean()),
                      "same_results_as_first": same}), flush=True)
# the start ball of the headline bench (65 536 walkers), default options; sha256 of the results for build-to-build comparisons
import hashlib
tb = torch.from_numpy(T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 65536, 1)).cuda()
lb = torch.empty(65536, dtype=torch.float64, device="cuda"); sb = torch.empty(65536, dtype=torch.int32, device="cuda")
m.set_option("dense_output", 0); m.set_option("cost_order", 1)
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.loglik_dev(oh, tb.data_ptr(), 65536, lb.data_ptr(), sb.data_ptr(), s); e1.record()
    torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print(json.dumps({"start_ball_walkers": 65536, "ms": best, "evals_per_s": 65536 / best * 1e3,
                  "sha256_logp_status": hashlib.sha256(lb.cpu().numpy().tobytes() + sb.cpu().numpy().tobytes()).hexdigest(),
                  "sha256_posterior_default": hashlib.sha256(ref[0].tobytes() + ref[1].tobytes()).hexdigest()}), flush=True)

