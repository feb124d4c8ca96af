#!/usr/bin/env python3
"""Long-run posterior agreement on the headline problem (HD155358.vels, 2 planets, 10 parameters, logp = -chi2/100 as the
reference defines it).

The affine stretch ensemble is run to stationarity from the reference's 1e-3 start ball.  SMALA with the reference's
eps = 0.025 mixes slowly (AC times 179-648 iterations, HD155358.ipynb:22341-22348; round-1 measurement: 6000 steps from
the ball are not enough to reach the stationary spread), so SMALA and MH are tested for INVARIANCE instead: their chains
start from walkers of the equilibrated ensemble and must keep its distribution -- a biased kernel would drift away from it
within a few autocorrelation times.  Reported per parameter: means, standard deviations, the mean difference in units of
its Monte-Carlo standard error and the two-sample KS distance the reference uses as its cross-sampler criterion
(driver.py:416-425).  One GPU, about 8 minutes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
KEYS = ["a0", "h0", "k0", "m0", "l0", "a1", "h1", "k1", "m1", "l1"]


def compare(ref, other, n_indep_other, driver):
    """ref[rows][walkers][10] (stationary ensemble), other[rows][chains][10]; n_indep_other: conservative count of
    independent samples in `other`."""
    rows = []
    for i, k in enumerate(KEYS):
        a, b = ref[:, :, i].ravel(), other[:, :, i].ravel()
        n_ref = ref.shape[1] * max(1, ref.shape[0] // 4)          # walkers x (rows / ~4 AC times apart)
        mcse = np.sqrt(a.var() / n_ref + b.var() / n_indep_other)
        ks = driver.calc_kstatistic(a[::7].reshape(-1, 1), b[::3].reshape(-1, 1))[0]
        rows.append({"param": k, "mean_ref": a.mean(), "mean": b.mean(), "std_ref": a.std(), "std": b.std(),
                     "diff_over_mcse": (b.mean() - a.mean()) / mcse, "ks": ks})
    return rows


def main():
    import rvtest as T
    from rvel_mcmc_b200 import _abi, driver
    ctx = _abi.Context(0)
    obs = T.load_vels("HD155358.vels")
    oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
    m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    n_st = int(os.environ.get("STRETCH_STEPS", "4000"))
    n_sm = int(os.environ.get("SMALA_STEPS", "1500"))
    n_mh = int(os.environ.get("MH_STEPS", "3000"))
    w_st, w_sm = 2048, 592
    out = {}
    t0 = time.perf_counter()
    st = m.stretch_run(oh, T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, w_st, 1), n_st, seed=101, thin=10)
    out["stretch"] = {"walkers": w_st, "steps": n_st, "seconds": time.perf_counter() - t0,
                      "accept": float(st["n_accept"].mean() / n_st)}
    ref = st["chain"][st["chain"].shape[0] // 2:]                  # second half: stationary (std stable from 1500 steps on)
    # stationarity of the reference itself: third quarter vs fourth quarter
    q = ref.shape[0] // 2
    out["stretch"]["self_ks_max"] = float(max(driver.calc_kstatistic(ref[:q, :, i].reshape(-1, 1)[::5], ref[q:, :, i].reshape(-1, 1)[::5])[0]
                                              for i in range(10)))
    start = st["theta"]
    rng = np.random.RandomState(7)
    # ---- SMALA from equilibrated starts (reference step size / softabs constant, (Ex)HD155358.ipynb:640)
    t0 = time.perf_counter()
    sm = m.smala_run(oh, start[rng.choice(w_st, w_sm, replace=False)], 0.025, 1.4, n_sm, seed=102, thin=10)
    out["smala"] = {"chains": w_sm, "steps": n_sm, "seconds": time.perf_counter() - t0,
                    "accept": float(sm["n_accept"].mean() / n_sm), "not_spd": int((sm["status"] == 9).sum()),
                    "params": compare(ref, sm["chain"][sm["chain"].shape[0] // 2:], w_sm, driver)}
    # ---- MH from equilibrated starts, proposal scale = 0.25 x the posterior standard deviations
    sc = ref.reshape(-1, 10).std(axis=0)
    t0 = time.perf_counter()
    mh = m.mh_run(oh, start, sc, 0.25, n_mh, seed=103, thin=10)
    out["mh"] = {"chains": w_st, "steps": n_mh, "seconds": time.perf_counter() - t0,
                 "accept": float(mh["n_accept"].mean() / n_mh),
                 "params": compare(ref, mh["chain"][mh["chain"].shape[0] // 2:], w_st, driver)}
    for k in ("smala", "mh"):
        out[k]["max_abs_diff_over_mcse"] = float(max(abs(r["diff_over_mcse"]) for r in out[k]["params"]))
        out[k]["max_ks"] = float(max(r["ks"] for r in out[k]["params"]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
