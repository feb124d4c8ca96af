#!/usr/bin/env python3
"""Python-3 equivalents of the reference's stand-alone benchmark scripts, on the GPU backend (no plotting):

  python bench_scripts/mcmc_benchmarks.py mh    [--niter 6000]      mcmc_benchmark_mh.py     (config lines 32-34, 52-54)
  python bench_scripts/mcmc_benchmarks.py smala [--niter 4200]      mcmc_benchmark_smala.py  (lines 35, 37, 53-54)
  python bench_scripts/mcmc_benchmarks.py emcee [--niter 25000]     mcmc_benchmark_emcee.py  (lines 33-34, 50-52)

The reference scripts cannot run as they are even on the reference (Python-2 syntax, mcmc.Smala(true_state, obs) without
eps/alp, missing TEST_3-2_COMPACT.vels), so these keep their workload definitions -- true states, observation
generators, scales, step sizes, iteration counts, loop shape (step_force / Ensemble.step), AC-time printout -- through the
same State / Observations / mcmc API.  `--fused` runs the same sampler through the one-call device loop (driver.run_*_gpu).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvel_mcmc_b200 import driver, mcmc, observations, state  # noqa: E402


def ac_report(chain, keys):
    for i, k in enumerate(keys):
        print("AC time {k}: {t}".format(k=k, t=driver.ac_time(chain[:, i])))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["mh", "smala", "emcee"])
    ap.add_argument("--niter", type=int, default=0)
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    np.random.seed(args.seed)
    t0 = time.perf_counter()
    if args.which == "mh":
        true_state = state.State(planets=[{"m": 1.2e-3, "a": 0.88, "h": 0.218, "k": 0.015, "l": 0.3},
                                          {"m": 2.1e-3, "a": 1.44 + 0.11, "h": 0.16, "k": 0.02, "l": 2.2}])
        obs = observations.FakeObservation(true_state, Npoints=200, error=1.5e-4, errorVar=2.5e-5, tmax=120.)
        scales = {"m": 1.e-3, "a": 0.3, "h": 0.5, "k": 0.5, "l": np.pi / 2.}
        Niter = args.niter or 6000
        if args.fused:
            bundle, _ = driver.run_mh_gpu("mh", Niter, true_state, obs, scales, 10.0e-3, nchains=1, seed=args.seed)
            chain = bundle.mcmc_chain
        else:
            mh = mcmc.Mh(true_state, obs)
            mh.set_scales(scales)
            mh.step_size = 10.0e-3
            chain = np.zeros((Niter, mh.state.Nvars))
            tries = 0
            for i in range(Niter):
                tries += mh.step_force()
                chain[i] = mh.state.get_params()
            print("Acceptance rate: %.3f%%" % (float(Niter) / tries * 100))
        keys = true_state.get_keys()
    elif args.which == "smala":
        true_state = state.State(planets=[{"m": 0.9e-3, "a": 0.226, "h": -0.06, "k": -0.015, "l": 1.3},
                                          {"m": 1.85e-3, "a": 0.3057, "h": -0.03, "k": -0.01, "l": 1.75}])
        obs = observations.FakeObservation(true_state, Npoints=60, error=1.5e-4, errorVar=2.5e-5, tmax=30.)
        eps, alp = 0.3, 1.4
        Niter = args.niter or 4200
        if args.fused:
            bundle, _ = driver.run_smala_gpu("smala", Niter, true_state, obs, eps, alp, nchains=1, seed=args.seed)
            chain = bundle.mcmc_chain
        else:
            smala = mcmc.Smala(true_state, obs, eps, alp)
            chain = np.zeros((Niter, smala.state.Nvars))
            tries = 0
            for i in range(Niter):
                tries += smala.step_force()
                chain[i] = smala.state.get_params()
            print("Acceptance rate: %.2f%%" % (float(Niter) / tries * 100))
        keys = true_state.get_keys()
    else:
        true_state = state.State(planets=[{"m": 0.94e-3, "a": 0.226, "h": -0.045, "k": -0.015, "l": 1.265},
                                          {"m": 1.965e-3, "a": 0.307, "h": -0.035, "k": -0.00, "l": 1.76}])
        obs = observations.FakeObservation(true_state, Npoints=200, error=1.5e-4, errorVar=2.5e-5, tmax=30.)
        scales = {"m": 1.5e-3, "a": 0.3, "h": 0.1, "k": 0.1, "l": np.pi / 2.}
        Nwalkers = 32
        Niter = args.niter or 25000
        if args.fused:
            bundle, _ = driver.run_emcee_gpu("emcee", Niter, true_state, obs, Nwalkers, scales, seed=args.seed)
            chain = bundle.mcmc_chain
        else:
            ens = mcmc.Ensemble(true_state, obs, scales=scales, nwalkers=Nwalkers)
            per = Niter // Nwalkers
            chain = np.zeros((per * Nwalkers, ens.state.Nvars))
            for i in range(per):
                ens.step()
                for j in range(Nwalkers):
                    chain[j * per + i] = ens.states[j]
        keys = true_state.get_keys()
    dt = time.perf_counter() - t0
    print("run '%s'%s: %d iterations in %.2f s (%.1f it/s)" % (args.which, " fused" if args.fused else "", len(chain), dt, len(chain) / dt))
    print("mean:", chain[len(chain) // 4:].mean(axis=0))
    print("true:", true_state.get_params())
    ac_report(chain[len(chain) // 4:], keys)


if __name__ == "__main__":
    main()
