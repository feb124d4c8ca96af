#!/usr/bin/env python3
"""Throughput of value + gradient + Hessian for models beyond two planets (thread per (set, planet); second-order sets in
several launches where one thread block does not hold them): C4 (three planets, 15 parameters), four and five planets.
Usage: python tools/time_var_many.py [walkers, default 296]"""
import json
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvtest as T
import parity_horizon as PH
from rvel_mcmc_b200 import _abi

W = int(sys.argv[1]) if len(sys.argv) > 1 else 296
ctx = _abi.Context(0)
s = torch.cuda.current_stream().cuda_stream
for name in ("c4", "four", "five"):
    obs, fixed, fp, fe, hill, center, sc = PH.problem(name)
    oh, m = PH._handles(ctx, obs, fixed, fp, fe, hill)
    nv = len(fp)
    theta = torch.from_numpy(T.gaussian_ball(center, sc, W, 1, width=1e-3)).cuda()
    lp = torch.empty(W, dtype=torch.float64, device="cuda"); st = torch.empty(W, dtype=torch.int32, device="cuda")
    g = torch.empty((W, nv), dtype=torch.float64, device="cuda"); h = torch.empty((W, nv, nv), dtype=torch.float64, device="cuda")
    m.loglik_d_dd_dev(oh, theta.data_ptr(), 8, lp.data_ptr(), g.data_ptr(), h.data_ptr(), st.data_ptr(), s); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.loglik_d_dd_dev(oh, theta.data_ptr(), W, lp.data_ptr(), g.data_ptr(), h.data_ptr(), st.data_ptr(), s); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nsets = 1 + nv + nv * (nv + 1) // 2
    print(json.dumps({"problem": name, "planets": int(fixed.shape[0]), "free_parameters": nv, "variational_sets": nsets,
                      "epochs": len(obs.tf) + len(obs.tb), "walkers": W, "ms": ms, "var_evals_per_s": W / ms * 1e3,
                      "ok": float((st.cpu().numpy() == 0).mean())}), flush=True)
    m.close(); oh.close()
