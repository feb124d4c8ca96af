#!/usr/bin/env python3
"""BASELINE.json configs on the GPU (one rank per GPU under torchrun, or a single process):

  C1  single-planet synthetic RV ("Simplest test"), Metropolis-Hastings                      (configs[0])
  C2  HD155358.vels two-planet fit, affine stretch sampler, 8 ... 65536 walkers              (configs[1])
  C3  HD155358 two-planet SMALA (second-order variational equations)                         (configs[2])
  C4  synthetic 3-planet near-resonant system, 10^5 independent MH chains (sharded by rank)   (configs[3])
  C5  throughput sweep: walkers x epochs, HD155358-shape truth, synthetic epochs              (configs[4])

Prints one JSON object per line (rank 0).  Timing: CUDA events / wall clock around synchronous ABI calls, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HD = [{"a": 0.65773033, "h": -0.0972263877, "k": -0.0782798396, "m": 0.000884031737, "l": 4.4280499},
      {"a": 1.04404207, "h": -0.0205622789, "k": -0.108797961, "m": 0.00083037971, "l": 1.49919861}]
HD_SCALES = {"m": 5.5e-6, "a": 0.001, "h": 0.02, "k": 0.02, "l": np.pi / 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rvel_mcmc_b200 import _abi, observations, state
    from rvel_mcmc_b200.samplers import chain_shard

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    ctx = _abi.Context(local)
    _abi.set_default_context(ctx)

    def maxsec(sec):
        if world == 1:
            return sec
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def emit(d):
        d["n_gpus"] = world
        if rank == 0:
            print(json.dumps(d), flush=True)

    def want(c):
        return not args.only or c in args.only.split(",")

    def scale_vec(st, scales):
        return np.array([scales[k] for k in st.get_rawkeys()])

    # ---- C1 -------------------------------------------------------------------------------------------------
    if want("C1"):
        np.random.seed(200000)
        true = state.State([{"a": 0.35, "m": 0.001965}], ignore_vars=["m"])          # Simplest test Long.ipynb:60-62
        obs = observations.FakeObservation(true, Npoints=100, error=3e-4, errorVar=9e-5, tmax=1.7)
        m, oh = true._model(ctx), obs._handle(ctx)
        W = 65536 if not args.quick else 4096
        lo, hi = chain_shard(W, rank, world)
        nsteps = 200
        th0 = np.tile(true.get_params(), (hi - lo, 1))
        m.mh_run(oh, th0[:64], [3e-4], 5.0, 2, seed=1)
        t0 = time.perf_counter()
        r = m.mh_run(oh, th0, [3e-4], 5.0, nsteps, seed=1, first_chain_id=lo, thin=4)
        sec = maxsec(time.perf_counter() - t0)
        emit({"config": "C1 single-planet synthetic (Simplest test), MH", "chains": W, "steps": nsteps, "epochs": 101,
              "evals_per_s": W * (nsteps + 1) / sec, "accept_rate": float(r["n_accept"].mean() / nsteps),
              "posterior_mean_a": float(r["chain"][10:].mean()), "true_a": 0.35})

    # ---- C2 -------------------------------------------------------------------------------------------------
    if want("C2"):
        obs = observations.Observation_FromFile(os.path.join(ROOT, "tests", "golden", "HD155358.vels"), Npoints=100)
        st = state.State([dict(p) for p in HD]); st.hillRadiusFactor = 2.
        m, oh = st._model(ctx), obs._handle(ctx)
        sc = scale_vec(st, HD_SCALES)
        from rvel_mcmc_b200.samplers import stretch_run_sharded
        for W in ([8, 64, 1024, 16384, 65536] if not args.quick else [8, 1024]):
            if (W // 2) % world:
                continue
            rng = np.random.RandomState(3)
            th0 = st.get_params()[None, :] + 1e-3 * sc[None, :] * rng.normal(size=(W, 10))
            nsteps = 20
            stretch_run_sharded(m, oh, th0, 1, seed=5, dist=dist if world > 1 else None, record_chain=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = stretch_run_sharded(m, oh, th0, nsteps, seed=5, dist=dist if world > 1 else None, record_chain=False)
            sec = maxsec(time.perf_counter() - t0)
            emit({"config": "C2 HD155358 two-planet, affine stretch", "walkers": W, "ensemble_steps": nsteps, "epochs": 122,
                  "evals_per_s": W * (nsteps + 1) / sec, "ms_per_ensemble_step": 1e3 * sec / nsteps,
                  "accept_rate": float(r["n_accept"].sum() / (W * nsteps)), "degenerate": bool(W < 11)})

    # ---- C3 -------------------------------------------------------------------------------------------------
    if want("C3"):
        obs = observations.Observation_FromFile(os.path.join(ROOT, "tests", "golden", "HD155358.vels"), Npoints=100)
        st = state.State([dict(p) for p in HD]); st.hillRadiusFactor = 2.
        m, oh = st._model(ctx), obs._handle(ctx)
        sc = scale_vec(st, HD_SCALES)
        W = 2368 if not args.quick else 296
        lo, hi = chain_shard(W * world, rank, world)
        rng = np.random.RandomState(4 + rank)
        th0 = st.get_params()[None, :] + 1e-3 * sc[None, :] * rng.normal(size=(hi - lo, 10))
        nsteps = 10
        m.smala_run(oh, th0[:8], 0.025, 1.4, 1, seed=2)
        t0 = time.perf_counter()
        r = m.smala_run(oh, th0, 0.025, 1.4, nsteps, seed=2, first_chain_id=lo)                # (Ex)HD155358.ipynb:640
        sec = maxsec(time.perf_counter() - t0)
        emit({"config": "C3 HD155358 two-planet SMALA (eps 0.025, alpha 1.4)", "chains": W * world, "steps": nsteps,
              "var_evals_per_s": W * world * (nsteps + 1) / sec, "accept_rate": float(r["n_accept"].mean() / nsteps),
              "not_spd_flags": int((r["status"] == 9).sum())})

    # ---- C4 -------------------------------------------------------------------------------------------------
    if want("C4"):
        np.random.seed(17)
        planets = [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0},             # mcmc_benchmark_smala.py:32 (2:1)
                   {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1},
                   {"m": 1.0e-3, "a": 0.59, "h": 0.0, "k": 0.03, "l": 0.7}]                    # third planet near the next 2:1
        true = state.State(planets)
        obs = observations.FakeObservation(true, Npoints=150, error=1.5e-4, errorVar=2.5e-5, tmax=60.)   # mcmc_benchmark_mh.py:34
        m, oh = true._model(ctx), obs._handle(ctx)
        from rvel_mcmc_b200.samplers import ess
        W = 100000 if not args.quick else 8192
        lo, hi = chain_shard(W, rank, world)
        th0 = np.tile(true.get_params(), (hi - lo, 1))

        def c4_row(label, sc, step, nsteps, thin):
            m.mh_run(oh, th0[:64], sc, step, 1, seed=3)
            t0 = time.perf_counter()
            r = m.mh_run(oh, th0, sc, step, nsteps, seed=3, first_chain_id=lo, thin=max(thin, 1), record_chain=thin > 0,
                         chain_walkers=256)
            sec = maxsec(time.perf_counter() - t0)
            row = {"config": "C4 synthetic 3-planet near-resonant, independent MH chains", "proposal": label, "chains": W,
                   "steps": nsteps, "epochs": 151, "nvars": 15, "step_size": step, "evals_per_s": W * (nsteps + 1) / sec,
                   "accept_rate": float(r["n_accept"].mean() / nsteps)}
            if thin > 0 and r["chain"] is not None and r["chain"].shape[0] >= 16:
                n_eff, tau = ess(r["chain"][r["chain"].shape[0] // 4:])          # drop the first quarter (start = truth)
                row["tau_int_max_rows"] = tau
                row["thin"] = thin
                # all chains are statistically identical: ESS of the whole run = rows x chains / tau, over the whole wall time
                row["ess_per_s_this_rank"] = (r["chain"].shape[0] * 3 // 4) * (hi - lo) / max(tau, 1.0) / sec
            emit(row)

        # (a) the reference's own proposal (mcmc_benchmark_mh.py:52-53): steps of 1 % in a and 5e-3 in h, k are hundreds of
        #     posterior widths wide, so essentially nothing is accepted -- kept as a labelled row, it measures rejected proposals
        sc_ref = scale_vec(true, {"m": 1e-3, "a": 0.3, "h": 0.5, "k": 0.5, "l": np.pi / 2})
        c4_row("reference scales, step 1e-2 (mcmc_benchmark_mh.py:52-53)", sc_ref, 1e-2, 20 if not args.quick else 5, 0)
        # (b) tuned: per-parameter scale = conditional posterior width 1/sqrt(-H_ii) from the variational Hessian at the truth
        #     (rv_loglik_d_dd), step 0.6 -> acceptance ~0.3
        _, _, hs, stt = m.loglik_d_dd(oh, true.get_params()[None, :])
        assert stt[0] == 0
        sc_tuned = 1.0 / np.sqrt(-np.diag(hs[0]))
        c4_row("tuned: scales 1/sqrt(-H_ii) at the truth, step 0.6", sc_tuned, 0.6, 200 if not args.quick else 40, 2)

    # ---- C5 -------------------------------------------------------------------------------------------------
    if want("C5"):
        st = state.State([dict(p) for p in HD]); st.hillRadiusFactor = 2.
        sc = scale_vec(st, HD_SCALES)
        p_inner = 2 * np.pi * HD[0]["a"] ** 1.5
        grid = [(10 ** 3, 100), (10 ** 4, 100), (10 ** 5, 100), (10 ** 6, 100), (10 ** 7, 100),
                (10 ** 5, 50), (10 ** 5, 200), (10 ** 5, 500), (10 ** 5, 1000)]
        if args.quick:
            grid = [(10 ** 4, 50), (10 ** 4, 200)]
        for W, nep in grid:
            np.random.seed(1000 + nep)
            obs = observations.FakeObservation(st, Npoints=nep, error=1.5e-4, errorVar=2.5e-5, tmax=10 * p_inner * nep / 100.)
            m, oh = st._model(ctx), obs._handle(ctx)
            lo, hi = chain_shard(W, rank, world)
            rng = np.random.RandomState(5 + rank)
            th = torch.from_numpy(st.get_params()[None, :] + 1e-3 * sc[None, :] * rng.normal(size=(hi - lo, 10))).cuda()
            lp = torch.empty(hi - lo, dtype=torch.float64, device="cuda"); stt = torch.empty(hi - lo, dtype=torch.int32, device="cuda")
            s = torch.cuda.current_stream().cuda_stream
            m.loglik_dev(oh, th.data_ptr(), min(hi - lo, 1024), lp.data_ptr(), stt.data_ptr(), s); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = None
            for _ in range(2):      # the first full-size call also grows the context's scratch buffers; report the second
                e0.record(); m.loglik_dev(oh, th.data_ptr(), hi - lo, lp.data_ptr(), stt.data_ptr(), s); e1.record()
                torch.cuda.synchronize()
                best = e0.elapsed_time(e1) * 1e-3
            sec = maxsec(best)
            emit({"config": "C5 sweep, HD155358-shape truth, IAS15", "walkers": W, "epochs": nep + 1, "ms": 1e3 * sec,
                  "evals_per_s": W / sec, "ok_fraction": float((stt == 0).float().mean().item())})
            if W == 10 ** 5 or args.quick:
                # the optional WHFast variant on the same batch, dt = P_inner/20 (BASELINE configs[4]; parity unpinned)
                ref = lp.clone()
                m.set_option("dt0", p_inner / 20.); m.set_option("integrator", 1)
                m.loglik_dev(oh, th.data_ptr(), min(hi - lo, 1024), lp.data_ptr(), stt.data_ptr(), s); torch.cuda.synchronize()
                e0.record(); m.loglik_dev(oh, th.data_ptr(), hi - lo, lp.data_ptr(), stt.data_ptr(), s); e1.record()
                torch.cuda.synchronize()
                sec = maxsec(e0.elapsed_time(e1) * 1e-3)
                okb = stt == 0
                emit({"config": "C5 sweep, HD155358-shape truth, WHFast dt=P_inner/20", "walkers": W, "epochs": nep + 1,
                      "ms": 1e3 * sec, "evals_per_s": W / sec, "ok_fraction": float(okb.float().mean().item()),
                      "max_abs_dlogp_vs_ias15": float((lp[okb] - ref[okb]).abs().max().item())})
                m.set_option("integrator", 0); m.set_option("dt0", 1e-3)
                # IAS15 with the non-default dense_output option (natural steps, RVs read inside the steps)
                m.set_option("dense_output", 1)
                m.loglik_dev(oh, th.data_ptr(), min(hi - lo, 1024), lp.data_ptr(), stt.data_ptr(), s); torch.cuda.synchronize()
                e0.record(); m.loglik_dev(oh, th.data_ptr(), hi - lo, lp.data_ptr(), stt.data_ptr(), s); e1.record()
                torch.cuda.synchronize()
                sec = maxsec(e0.elapsed_time(e1) * 1e-3)
                okb = stt == 0
                emit({"config": "C5 sweep, HD155358-shape truth, IAS15 dense_output=1", "walkers": W, "epochs": nep + 1,
                      "ms": 1e3 * sec, "evals_per_s": W / sec, "ok_fraction": float(okb.float().mean().item()),
                      "max_abs_dlogp_vs_default": float((lp[okb] - ref[okb]).abs().max().item())})
                m.set_option("dense_output", 0)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
