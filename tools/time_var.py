import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__))+"/.."); sys.path.insert(0, ""+__import__("os").path.dirname(__import__("os").path.abspath(__file__))+"/../tests")
import rvtest as T
from rvel_mcmc_b200 import _abi
ctx = _abi.Context(0)
obs = T.load_vels("HD155358.vels")
oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
for W in (296, 2368, 8192):
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 1)
    m.loglik_d_dd(oh, theta[:8])
    t0 = time.perf_counter(); r = m.loglik_d_dd(oh, theta); dt = time.perf_counter() - t0
    print("W=%d  %.3f s  %.1f var-evals/s  ok=%.3f" % (W, dt, W / dt, (r[3] == 0).mean()))
