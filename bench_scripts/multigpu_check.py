#!/usr/bin/env python3
"""Multi-GPU check + timing of the sharded samplers (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      bench_scripts/multigpu_check.py [--walkers 4096] [--steps 8]

1. affine stretch ensemble sharded over N ranks (NCCL all-gather of the updated half after each half-step) must equal
   the single-GPU rv_stretch_run bit for bit;
2. MH chains sharded by contiguous blocks (no collective) must equal the single-GPU chains bit for bit;
3. prints evaluations/s of the sharded stretch ensemble (device time, max over ranks).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--ess-steps", type=int, default=0, help="also run this many recorded ensemble steps and report ESS/s")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import rvtest as T
    from rvel_mcmc_b200 import _abi
    from rvel_mcmc_b200.samplers import stretch_run_sharded, chain_shard

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    ctx = _abi.Context(local)
    obs = T.load_vels("HD155358.vels")
    oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
    m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    W, nsteps = args.walkers, args.steps
    theta0 = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 21)
    d = dist if world > 1 else None
    stretch_run_sharded(m, oh, theta0[: 2 * world * 8], 1, seed=1, dist=d)        # warm-up
    torch.cuda.synchronize()
    if d: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = stretch_run_sharded(m, oh, theta0, nsteps, seed=77, dist=d, record_chain=False)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if d: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"n_gpus": world, "walkers": W, "ensemble_steps": nsteps, "ms": float(ms.item()),
           "stretch_evals_per_s": W * (nsteps + 1) / (float(ms.item()) * 1e-3)}
    if args.ess_steps:
        from rvel_mcmc_b200.samplers import ess
        torch.cuda.synchronize()
        if d: dist.barrier()
        t0 = time.perf_counter()
        rr = stretch_run_sharded(m, oh, theta0, args.ess_steps, seed=78, dist=d, record_chain=True, thin=1)
        sec = time.perf_counter() - t0
        if rank == 0:
            c = rr["chain"][args.ess_steps // 4:, ::max(1, W // 256), :]          # tau from a 256-walker subsample
            n_eff, tau = ess(c)
            out["ess"] = {"ensemble_steps": args.ess_steps, "seconds": sec, "tau_int_max": tau,
                          "ess_per_s": (args.ess_steps - args.ess_steps // 4) * W / max(tau, 1.0) / sec,
                          "evals_per_s": W * (args.ess_steps + 1) / sec}
    # reference results on one GPU (rank 0 only)
    lo, hi = chain_shard(W, rank, world)
    scales = np.array(T.HD_SCALE_VEC)
    mh_loc = m.mh_run(oh, theta0[lo:hi], scales, 0.5, 6, seed=9, first_chain_id=lo, record_chain=False, record_accepts=True)
    if rank == 0:
        single = m.stretch_run(oh, theta0, nsteps, seed=77, record_chain=False)
        out["stretch_equal_single_gpu"] = bool(np.array_equal(single["theta"], r["theta"]) and np.array_equal(single["lnp"], r["lnp"]))
        out["stretch_accept_rate"] = float(single["n_accept"].sum() / (W * nsteps))
        mh_all = m.mh_run(oh, theta0, scales, 0.5, 6, seed=9, record_chain=False, record_accepts=True)
        out["mh_shard_equal_single_gpu"] = bool(np.array_equal(mh_all["theta"][lo:hi], mh_loc["theta"]) and
                                                np.array_equal(mh_all["accepted"][:, lo:hi], mh_loc["accepted"]))
        print(json.dumps(out))
        ok = out["stretch_equal_single_gpu"] and out["mh_shard_equal_single_gpu"]
    else:
        ok = True
    if d:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
