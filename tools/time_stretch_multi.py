#!/usr/bin/env python3
"""Strong scaling of the affine-stretch ensemble through the single-process C-ABI path (rv_stretch_run_multi: every GPU
holds a full copy of the ensemble, the accept kernel of a slice stores its accepted walkers into all copies over peer
memory) against the single-GPU call (rv_stretch_run) on the same ensemble, same seed.

Timing: the calls are synchronous and take host buffers, so each is timed by the host clock for two step counts and the
per-step time is the DIFFERENCE (upload, download and start-up cancel); the walkers are the committed equilibrated HD155358
ensemble tiled to the requested size.  Prints one JSON line per configuration.
Usage: python tools/time_stretch_multi.py [walkers, default 65536] [steps, default 12]"""
import json
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rvtest as T
from rvel_mcmc_b200 import observations, state, _abi
from rvel_mcmc_b200.multigpu import DeviceGroup

W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n1 = max(1, n2 // 4)
obs = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
st = state.State(T.planets_from_vec(T.HD_SOL)); st.hillRadiusFactor = 2.
ens = np.load(os.path.join(ROOT, "tests", "golden", "hd155358_equilibrated_ensemble.npy"))
rng = np.random.RandomState(5)
theta = np.ascontiguousarray(np.tile(ens, (W // len(ens) + 1, 1))[rng.permutation((W // len(ens) + 1) * len(ens))[:W]])
# tiled copies are identical walkers: the stretch move between two equal walkers is a no-op, so jitter them inside the posterior
theta = theta * (1.0 + 1e-6 * rng.standard_normal(theta.shape))

g = DeviceGroup()
G = len(g)
ctx0 = g.ctxs[0]
m0, o0 = st._model(ctx0), obs._handle(ctx0)


def timed(fn, n):
    t0 = time.perf_counter(); r = fn(n); return time.perf_counter() - t0, r


raw = {}


def per_step(fn, name):
    fn(1)                                            # warm-up: allocations, peer mappings, clocks
    ta = min(timed(fn, n1)[0] for _ in range(2))
    tb, r = timed(fn, n2)
    tb2, r = timed(fn, n2)
    raw[name] = {"s_for_%d_steps" % n1: ta, "s_for_%d_steps" % n2: [tb, tb2]}
    return (min(tb, tb2) - ta) / (n2 - n1), r


single = lambda n: m0.stretch_run(o0, theta, n, seed=11, record_chain=False)
multi = lambda n: g.stretch_run(st, obs, theta, n, seed=11)
s1, r1 = per_step(single, "single_gpu")
out = {"walkers": W, "gpus": G, "steps": [n1, n2], "timing": "host clock, difference of two step counts",
       "single_gpu": {"ms_per_ensemble_step": 1e3 * s1, "evals_per_s": W / s1}}
if G > 1 and (W // 2) % G == 0:
    sm, rm = per_step(multi, "multi")
    out["multi"] = {"call": "rv_stretch_run_multi (peer-store accept kernel)", "ms_per_ensemble_step": 1e3 * sm,
                    "evals_per_s": W / sm, "speedup_vs_single_gpu": s1 / sm,
                    "bit_identical_to_single_gpu": bool(np.array_equal(rm["theta"], r1["theta"]) and np.array_equal(rm["lnp"], r1["lnp"])
                                                        and np.array_equal(rm["n_accept"], r1["n_accept"]))}
    # weak: G times the walkers on G GPUs against W on one
    thw = np.ascontiguousarray(np.tile(theta, (G, 1)) * (1.0 + 1e-6 * rng.standard_normal((G * W, theta.shape[1]))))
    multi_w = lambda n: g.stretch_run(st, obs, thw, n, seed=11)
    sw, _ = per_step(multi_w, "multi_weak")
    out["multi_weak"] = {"walkers": G * W, "ms_per_ensemble_step": 1e3 * sw, "evals_per_s": G * W / sw,
                         "speedup_vs_single_gpu": (G * W / sw) / (W / s1)}
out["raw_seconds"] = raw
print(json.dumps(out), flush=True)
g.close()
