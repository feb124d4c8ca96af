#!/usr/bin/env python3
"""The reference's "(Ex)Full Test + Usage Example" notebook as a script, on the GPU backend (no plots).

One planet {a, h, k} free, m fixed; 70 synthetic epochs (error 3.5e-4, errorVar 9e-5, tmax 1.37), seed 200000; the four
samplers with the notebook's settings; acceptance, AC times (first lag with autocorrelation < 0.5), efficacy
Niter/(dt * max AC) and the cross-sampler KS distances.  Two ways of running each sampler are shown:

  step-by-step   driver.run_mh / run_emcee / run_smala / run_alsmala -- the reference's loops, one Python iteration and one
                 library call per step (bound by the latency of one integration on the GPU);
  fused          driver.run_*_gpu -- the whole loop of many chains in one library call.

  python examples/usage_example.py [--niter 2000] [--chains 256]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rvel_mcmc_b200 import driver, observations, state  # noqa: E402


def report(name, bundle, seconds, niter):
    act = driver.ac_times(bundle)
    print("%-22s %8d iterations  %7.2f s  %9.1f it/s   AC times %s   efficacy %.1f /s"
          % (name, niter, seconds, niter / seconds, np.round(act, 1), niter / (seconds * max(act.max(), 1.0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--niter", type=int, default=2000)
    ap.add_argument("--chains", type=int, default=256)
    a = ap.parse_args()
    np.random.seed(200000)
    true_state = state.State([{"a": 0.2275, "h": 0., "k": 0., "m": 0.001965}], ignore_vars=["m"])
    obs = observations.FakeObservation(true_state, Npoints=70, error=3.5e-4, errorVar=9e-5, tmax=1.37)
    scal = {'a': 3e-4, 'h': 0.01, 'k': 0.01}
    print("true parameters", dict(zip(true_state.get_rawkeys(), true_state.get_params())), " logp(true) = %.6f" % true_state.get_logp(obs))

    runs = {}
    for name, fn in (("mh step-by-step", lambda: driver.run_mh("ex", a.niter, true_state, obs, scal, 5, printing_every=10 ** 9)),
                     ("emcee step-by-step", lambda: driver.run_emcee("ex", a.niter, true_state, obs, 32, scal, printing_every=10 ** 9)),
                     ("smala step-by-step", lambda: driver.run_smala("ex", a.niter // 4, true_state, obs, 1.2, 0.14, printing_every=10 ** 9)),
                     ("alsmala step-by-step", lambda: driver.run_alsmala("ex", a.niter // 4, true_state, obs, 1.2, 0.14, 3.0, 0.0, printing_every=10 ** 9))):
        t0 = time.perf_counter()
        bundle, _ = fn()
        runs[name] = bundle
        report(name, bundle, time.perf_counter() - t0, bundle.mcmc_Niter)
    c = a.chains
    for name, fn in (("mh fused", lambda: driver.run_mh_gpu("ex", a.niter, true_state, obs, scal, 5, nchains=c, seed=1)),
                     ("emcee fused", lambda: driver.run_emcee_gpu("ex", a.niter * c, true_state, obs, c, scal, seed=2)),
                     ("smala fused", lambda: driver.run_smala_gpu("ex", a.niter // 4, true_state, obs, 1.2, 0.14, nchains=c, seed=3)),
                     ("alsmala fused", lambda: driver.run_alsmala_gpu("ex", a.niter // 4, true_state, obs, 1.2, 0.14, 3.0, nchains=c, seed=4))):
        t0 = time.perf_counter()
        bundle, _ = fn()
        runs[name] = bundle
        report(name, bundle, time.perf_counter() - t0, bundle.mcmc_Niter)

    def tail(b):
        ch = b.mcmc_chain
        if b.mcmc_is_emcee:
            per = ch.shape[0] // b.mcmc_Nwalkers
            return ch.reshape(b.mcmc_Nwalkers, per, -1)[:, per // 4:, :].reshape(-1, ch.shape[1])
        return ch[len(ch) // 4:]
    ref = tail(runs["mh fused"])
    print("posterior mean (mh fused):", ref.mean(axis=0), " std:", ref.std(axis=0))
    for name in ("emcee fused", "smala fused", "alsmala fused", "mh step-by-step", "smala step-by-step"):
        ks = driver.calc_kstatistic(ref[::5], tail(runs[name]))
        print("KS distance mh fused vs %-20s %s" % (name, np.round(ks, 3)))


if __name__ == "__main__":
    main()
