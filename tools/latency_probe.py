#!/usr/bin/env python3
"""Small-batch latency of the plain log-likelihood (BASELINE configs[0]/[1] low end): one State.get_logp-sized call and
small stretch ensembles, under the default step sequence and the two non-default epoch-handling options.
Prints one JSON line per measurement."""
import json
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvtest as T
from rvel_mcmc_b200 import _abi

ctx = _abi.Context(0)
obs = T.load_vels("HD155358.vels")
oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
for opts in ({}, {"monotone_backward": 1}, {"dense_output": 1}):
    m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    for k, v in opts.items():
        m.set_option(k, v)
    for W in (1, 8, 64, 1024):
        theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 1)
        m.loglik(oh, theta)
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            m.loglik(oh, theta)
        dt = (time.perf_counter() - t0) / n
        print(json.dumps({"what": "rv_loglik host call", "options": opts, "walkers": W, "ms_per_call": 1e3 * dt, "evals_per_s": W / dt}))
    for W in (8, 64, 1024):
        theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 2)
        m.stretch_run(oh, theta, 2, seed=1, record_chain=False)
        n = 20
        t0 = time.perf_counter()
        m.stretch_run(oh, theta, n, seed=1, record_chain=False)
        dt = time.perf_counter() - t0
        print(json.dumps({"what": "rv_stretch_run", "options": opts, "walkers": W, "ms_per_ensemble_step": 1e3 * dt / n,
                          "evals_per_s": W * (n + 1) / dt}))
    m.close()
