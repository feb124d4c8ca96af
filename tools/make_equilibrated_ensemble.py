#!/usr/bin/env python3
"""Generates tests/golden/hd155358_equilibrated_ensemble.npy -- the start ensemble of bench.py's ESS/s block.

The affine stretch ensemble (rv_stretch_run, emcee's move with a = 2) is run on the GPU from the reference's 1e-3 start
ball (mcmc.py:49-51, scales (Ex)HD155358.ipynb:456) on HD155358.vels until it is stationary: the run is cut into quarters
and the two-sample KS distance between the third and the fourth quarter (the reference's cross-sampler criterion,
driver.py:416-425) must be below 0.03 for every parameter, while the first quarter -- still spreading from the ball -- must
differ.  The final positions are the fixture.  Chains started in the ball need ~2000 ensemble steps to reach the posterior's
spread (profiles/r01q_*), which is why bench.py does not burn in inside its clock.

  python tools/make_equilibrated_ensemble.py [walkers=28416] [steps=4000]      (one B200, a few minutes)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvtest as T
from rvel_mcmc_b200 import _abi, driver

W = int(sys.argv[1]) if len(sys.argv) > 1 else 28416
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
out_dir = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out")
ctx = _abi.Context(0)
obs = T.load_vels("HD155358.vels")
oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 20261018)
lnp = None
quarters = []
t0 = time.perf_counter()
acc = 0.0
for q in range(4):
    r = m.stretch_run(oh, theta, N // 4, seed=424242, first_step=q * (N // 4), lnp=lnp, record_chain=False)
    theta, lnp = r["theta"], r["lnp"]
    acc += float(r["n_accept"].mean())
    quarters.append(theta.copy())
    print("quarter %d done after %.1f s, std %s" % (q + 1, time.perf_counter() - t0, np.array2string(theta.std(axis=0), precision=3)), flush=True)
sec = time.perf_counter() - t0
ks34 = driver.calc_kstatistic(quarters[2], quarters[3])
ks14 = driver.calc_kstatistic(quarters[0], quarters[3])
info = {"walkers": W, "ensemble_steps": N, "seconds": sec, "accept_rate": acc / N, "ks_quarter3_vs_4": ks34, "ks_quarter1_vs_4": ks14,
        "mean": theta.mean(axis=0).tolist(), "std": theta.std(axis=0).tolist(), "finite_lnp_fraction": float(np.isfinite(lnp).mean())}
print(json.dumps(info))
assert max(ks34) < 0.03, ks34
os.makedirs(out_dir, exist_ok=True)
np.save(os.path.join(out_dir, "hd155358_equilibrated_ensemble.npy"), theta)
with open(os.path.join(out_dir, "hd155358_equilibrated_ensemble.json"), "w") as f:
    json.dump(info, f)
