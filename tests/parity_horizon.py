"""How long do the device samplers make IDENTICAL accept/reject decisions to the CPU oracle?

north_star: "with fixed RNG streams, MH and affine chains must show identical accept/reject decisions over the first
10^4 steps".  Both sides draw from the same counter-based Philox streams; for the stretch move the proposal arithmetic
itself (zz, q = fma(-zz, c - s, c)) is a fixed sequence of correctly-rounded operations shared by kernel and oracle, so
walker positions are bit-identical functions of the decision history and the two likelihood implementations (which
differ by rounding, ~1e-11 in logp) can only disagree when |lnpdiff - ln u| falls inside that rounding band.

Used by tests/test_gpu_parity_horizon.py (asserts) and tools/parity_horizon_report.py (writes the JSON summary
committed under profiles/).  The oracle is the checker; the GPU run goes through the C ABI.
"""
import ctypes as C
import os
import threading
import time

import numpy as np

import rvtest as T


def _first_divergence(acc_gpu, acc_orc):
    """First step at which any walker's decision differs (None: identical over the whole run)."""
    d = np.nonzero((acc_gpu != acc_orc).any(axis=1))[0]
    return None if len(d) == 0 else int(d[0])


def _both(gpu_fn, orc_fn):
    """Run the device sampler and the oracle concurrently (both release the GIL); returns (gpu, orc, seconds)."""
    out = {}

    def run(name, fn):
        t0 = time.perf_counter()
        out[name] = fn()
        out[name + "_s"] = time.perf_counter() - t0
    th = threading.Thread(target=run, args=("orc", orc_fn))
    th.start()
    run("gpu", gpu_fn)
    th.join()
    return out["gpu"], out["orc"], {"gpu_s": out["gpu_s"], "oracle_s": out["orc_s"]}


def problem(name):
    """(obs, fixed, fp, fe, hill, center, scale_vec) of a named parity problem."""
    if name == "small":          # configs[0] shape: one planet, free (a, h, k)
        from test_samplers_cpu import _small_problem
        obs, E, fp, fe, center = _small_problem()
        return obs, E, fp, fe, 1.0, center, np.array([3e-4, 0.01, 0.01])
    if name == "hd155358":       # configs[1] shape
        return T.load_vels("HD155358.vels"), np.zeros((2, 7)), T.FP10, T.FE10, 1.0, np.array(T.HD_SOL), np.array(T.HD_SCALE_VEC)
    if name == "c4":             # configs[3] shape: three planets, 15 free parameters
        obs, fixed, center, sc = T.c4_problem()
        return obs, fixed, T.FP15, T.FE15, 2.0, center, sc
    if name in ("four", "five"):  # beyond the BASELINE configs: four / five planets, 20 / 25 free parameters
        obs, fixed, fp, fe, center, sc = T.many_planet_problem(4 if name == "four" else 5)
        # (hill factor 1 for five planets: twice the outermost planet's Hill radius exceeds the spacing of the inner pair)
        return obs, fixed, fp, fe, (2.0 if name == "four" else 1.0), center, sc
    raise KeyError(name)


def _handles(ctx, obs, fixed, fp, fe, hill):
    from rvel_mcmc_b200 import _abi
    oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
    return oh, _abi.ModelHandle(ctx, fixed, fp, fe, hill)


def mh_horizon(ctx, name, W, nsteps, scales, step_size, seed, theta0=None, nthreads=None):
    obs, fixed, fp, fe, hill, center, sc = problem(name)
    oh, m = _handles(ctx, obs, fixed, fp, fe, hill)
    nthreads = nthreads or os.cpu_count() or 1
    theta0 = np.tile(center, (W, 1)) if theta0 is None else theta0
    scales = np.ascontiguousarray(scales, dtype=np.float64)

    def orc():
        th = np.ascontiguousarray(theta0.copy()); lp = np.zeros(W)
        acc = np.zeros((nsteps, W), dtype=np.uint8)
        fpa = np.array(fp, dtype=np.int32); fea = np.array(fe, dtype=np.int32)
        T.oracle().orc_mh_run(fixed.shape[0], T.vp(fixed), len(fp), T.vp(fpa), T.vp(fea), C.c_double(hill),
                                   T.vp(obs.tf), T.vp(obs.rvf), T.vp(obs.errorf), len(obs.tf),
                                   T.vp(obs.tb), T.vp(obs.rvb), T.vp(obs.errorb), len(obs.tb), C.c_double(obs.Npoints),
                                   T.vp(th), T.vp(lp), T.vp(scales), C.c_double(step_size), C.c_uint64(seed), C.c_uint64(0),
                                   0, nsteps, C.c_long(W), None, T.vp(acc), nthreads)
        return th, lp, acc

    r, (th_o, lp_o, acc_o), secs = _both(
        lambda: m.mh_run(oh, theta0, scales, step_size, nsteps, seed=seed, record_chain=False, record_accepts=True), orc)
    m.close(); oh.close()
    fin = np.isfinite(lp_o)
    return dict(sampler="mh", problem=name, chains=W, steps=nsteps, decisions=int(W * nsteps),
                first_divergent_step=_first_divergence(r["accepted"], acc_o),
                mismatched_decisions=int((r["accepted"] != acc_o).sum()), accept_rate=float(acc_o.mean()),
                max_abs_theta_diff=float(np.abs(r["theta"] - th_o).max()),
                max_abs_logp_diff=float(np.abs(r["logp"][fin] - lp_o[fin]).max()), **secs)


def stretch_horizon(ctx, name, W, nsteps, seed, width=1e-3, ball_seed=6, nthreads=None):
    obs, fixed, fp, fe, hill, center, sc = problem(name)
    oh, m = _handles(ctx, obs, fixed, fp, fe, hill)
    nthreads = nthreads or os.cpu_count() or 1
    theta0 = T.gaussian_ball(center, sc, W, ball_seed, width=width)

    def orc():
        th = np.ascontiguousarray(theta0.copy()); lnp = np.zeros(W)
        acc = np.zeros((nsteps, W), dtype=np.uint8)
        fpa = np.array(fp, dtype=np.int32); fea = np.array(fe, dtype=np.int32)
        T.oracle().orc_stretch_run(fixed.shape[0], T.vp(fixed), len(fp), T.vp(fpa), T.vp(fea), C.c_double(hill),
                                        T.vp(obs.tf), T.vp(obs.rvf), T.vp(obs.errorf), len(obs.tf),
                                        T.vp(obs.tb), T.vp(obs.rvb), T.vp(obs.errorb), len(obs.tb), C.c_double(obs.Npoints),
                                        T.vp(th), T.vp(lnp), 0, C.c_double(2.0), C.c_uint64(seed), 0, nsteps, C.c_long(W),
                                        None, T.vp(acc), nthreads)
        return th, lnp, acc

    r, (th_o, lnp_o, acc_o), secs = _both(
        lambda: m.stretch_run(oh, theta0, nsteps, a=2.0, seed=seed, record_chain=False, record_accepts=True), orc)
    m.close(); oh.close()
    fin = np.isfinite(lnp_o)
    return dict(sampler="stretch", problem=name, walkers=W, ensemble_steps=nsteps, decisions=int(W * nsteps),
                first_divergent_step=_first_divergence(r["accepted"], acc_o),
                mismatched_decisions=int((r["accepted"] != acc_o).sum()), accept_rate=float(acc_o.mean()),
                positions_bit_identical=bool(np.array_equal(r["theta"], th_o)),
                max_abs_theta_diff=float(np.abs(r["theta"] - th_o).max()),
                max_abs_lnp_diff=float(np.abs(r["lnp"][fin] - lnp_o[fin]).max()), **secs)
