#!/usr/bin/env python3
"""Small run of every kernel family (seconds), e.g. for compute-sanitizer where it is available (it is closed on the
build pool, so round 1 relied on the sequential host emulation of the block algorithms and oracle comparisons instead):
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvtest as T
from rvel_mcmc_b200 import _abi

ctx = _abi.Context(0)
obs = T.load_vels("HD155358.vels")
# a short observation set keeps the sanitizer run short
o = T.Obs()
o.tf = obs.tf[:6]; o.rvf = obs.rvf[:6]; o.errorf = obs.errorf[:6]
o.tb = obs.tb[-6:]; o.rvb = obs.rvb[-6:]; o.errorb = obs.errorb[-6:]; o.Npoints = 12
oh = _abi.ObsHandle(ctx, o.tf, o.rvf, o.errorf, o.tb, o.rvb, o.errorb, o.Npoints)
m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 40, 1)
theta[3, 3] = 1e-6
lg, sg = m.loglik(oh, theta)
rv, st = m.rv_curve(theta[:5], o.tf)
ic, st = m.initial_conditions(theta[:5])
lv, gv, hv, sv = m.loglik_d_dd(oh, theta[:6])
r = m.mh_run(oh, theta[:8], np.array(T.HD_SCALE_VEC), 0.1, 3, seed=1)
r = m.stretch_run(oh, theta[:8], 2, seed=1)
r = m.smala_run(oh, theta[:4], 0.025, 1.4, 2, seed=1)
r = m.alsmala_run(oh, theta[:4], 0.025, 1.4, 3.0, 3, seed=1)
m.set_option("integrator", 1); m.set_option("dt0", 0.1)
lw, sw = m.loglik(oh, theta)
print("ok", np.isfinite(lg[sg == 0]).all(), (sv == 0).all(), np.isfinite(lw[sw == 0]).all())
