#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json: RV log-lik evals/sec, HD155358 2-planet).

One "step" = one pass of the hot path over one batch of synthetic walkers: the log-likelihood
(State.get_logp, state.py:103) of `--walkers` parameter vectors per GPU drawn as the reference's ensemble
start ball (mcmc.py:49-51) around the published HD155358 solution, on the real HD155358.vels epochs
(122 epochs, Npoints=100, hillRadiusFactor=2).  Weak scaling: every rank owns its own walker shard, no
data-path collective.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--walkers 65536] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = evaluations/s with inputs resident in HBM (CUDA events on the
launching stream, max over ranks); `e2e` = the same through the host-buffer C-ABI call rv_loglik (pinned
host memory, H2D + D2H inside the timed region).  `--impl reference` times the reference's CPU path:
the oracle port of rebound's algorithm (oracle/, OpenMP over all host cores) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "rv_loglik_evals_per_sec_hd155358_2planet"
UNIT = "evals/s"
HD_SOL = [6.57730330e-01, -9.72263877e-02, -7.82798396e-02, 8.84031737e-04, 4.42804990e+00,
          1.04404207e+00, -2.05622789e-02, -1.08797961e-01, 8.30379710e-04, 1.49919861e+00]
HD_SCALES = {"m": 5.5e-6, "a": 0.001, "h": 0.02, "k": 0.02, "l": np.pi / 4}
FP10 = [0, 0, 0, 0, 0, 1, 1, 1, 1, 1]
FE10 = [1, 2, 3, 0, 4, 1, 2, 3, 0, 4]


def walker_ball(W, seed):
    rng = np.random.RandomState(seed)
    sc = np.array([HD_SCALES[k] for k in ("a", "h", "k", "m", "l")] * 2)
    return np.asarray(HD_SOL)[None, :] + 1e-3 * sc[None, :] * rng.normal(size=(W, 10))


def load_obs():
    from rvel_mcmc_b200 import observations
    return observations.Observation_FromFile(os.path.join(ROOT, "tests", "golden", "HD155358.vels"), Npoints=100)


def flops_per_eval(S, T):
    """W_eval = S*[3N*36 + N(N-1)*18] + T*3N*150 with N = 3 bodies."""
    return S * (9 * 36 + 6 * 18) + T * 9 * 150


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.samples = []
        self.reasons = set()
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def oracle_batch(obs, theta, nthreads):
    import rvtest as T
    o = T.Obs()
    o.tf, o.tb, o.rvf, o.rvb, o.errorf, o.errorb, o.Npoints = obs.tf, obs.tb, obs.rvf, obs.rvb, obs.errorf, obs.errorb, obs.Npoints
    t0 = time.perf_counter()
    logp, st, cnt = T.orc_logp_batch(np.zeros((2, 7)), FP10, FE10, 2.0, o, theta, nthreads=nthreads, lib=T.oracle_fast())
    return time.perf_counter() - t0, logp, st, cnt


def run_ess(ctx, model, oh):
    """ESS/s of the three device samplers on the HD155358 posterior (bounded runs; reference AC definition + Sokal tau)."""
    from rvel_mcmc_b200.samplers import ess
    from rvel_mcmc_b200 import driver
    out = {}
    sc = np.array([HD_SCALES[k] for k in ("a", "h", "k", "m", "l")] * 2)

    def summarise(name, chain, seconds, evals, burn):
        c = chain[burn:]
        n_eff, tau = ess(c)
        ac_ref = max(driver.ac_time(c[:, 0, i]) for i in range(c.shape[2]))     # driver.py:366-377 (first lag with AC < 0.5)
        # ESS of the post-burn-in samples divided by the WHOLE run time (burn-in included); tau in recorded rows
        out[name] = {"walkers": int(chain.shape[1]), "recorded_rows": int(chain.shape[0]), "seconds": seconds,
                     "evals_per_s": evals / seconds, "tau_int_max_rows": tau, "ac_time_ref_max_rows": ac_ref,
                     "ess_per_s": n_eff / seconds}

    # walker counts that fill the machine: 3 CTAs x 64 lane groups per SM, one (walker, leg) item per group
    slots = ctx.device_info()["sm_count"] * 3 * 64
    W, n = 2 * slots, 300                   # each half-ensemble = one backward + one forward round
    t0 = time.perf_counter()
    r = model.stretch_run(oh, walker_ball(W, 5), n, seed=11, thin=2)
    summarise("stretch", r["chain"], time.perf_counter() - t0, W * (n + 1), n // 8)
    out["stretch"]["thin"] = 2
    W, n = slots, 300
    t0 = time.perf_counter()
    r = model.mh_run(oh, walker_ball(W, 6), sc, 0.1, n, seed=12, thin=2)
    summarise("mh", r["chain"], time.perf_counter() - t0, W * (n + 1), n // 8)
    out["mh"]["thin"] = 2
    W, n = 2368, 60
    t0 = time.perf_counter()
    r = model.smala_run(oh, walker_ball(W, 7), 0.025, 1.4, n, seed=13, thin=1)
    summarise("smala", r["chain"], time.perf_counter() - t0, W * (n + 1), n // 4)
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; rebound itself is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    obs = load_obs()
    cores = os.cpu_count() or 1
    sample = 4096                      # ~2 s of work per step on 16 cores
    theta = walker_ball(sample, 1234)
    for _ in range(min(args.warmup, 1)):
        oracle_batch(obs, theta[: cores * 4], cores)
    times = []
    cnt = None
    for k in range(args.steps):
        dt, _, _, cnt = oracle_batch(obs, walker_ball(sample, 1234 + k), cores)
        times.append(dt)
    tot = float(np.sum(times))
    value = sample * args.steps / tot
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic walkers on HD155358.vels epochs",
            "config": {"workload": "HD155358 2-planet log-likelihood, %d-walker sample per step (bounded CPU sample)" % sample,
                       "walkers_per_step": sample, "epochs": 122},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d walkers x %d steps, oracle/rv_oracle.c (gcc -O3 -march=native, OpenMP, %d threads)" % (sample, args.steps, cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--walkers", type=int, default=65536, help="walkers per GPU per step")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mapping", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="walkers in the timed CPU sample (0: about 15 s of work)")
    ap.add_argument("--ess", action="store_true", help="also run the stretch / MH / SMALA samplers and report ESS/s")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from rvel_mcmc_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")        # keep stdout to the one JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    ctx = _abi.Context(local)
    obs = load_obs()
    oh = obs._handle(ctx)
    model = _abi.ModelHandle(ctx, np.zeros((2, 7)), FP10, FE10, 2.0)
    if args.mapping:
        model.set_option("mapping", args.mapping)
    W = args.walkers
    nsets = args.steps + args.warmup
    # a fresh batch of walkers for every step (global walker ids are disjoint across ranks)
    host_theta = [torch.from_numpy(walker_ball(W, 1000 * (rank + 1) + s)).pin_memory() for s in range(nsets)]
    d_theta = [t.to(dev, non_blocking=True) for t in host_theta]
    d_logp = torch.empty(W, dtype=torch.float64, device=dev)
    d_status = torch.empty(W, dtype=torch.int32, device=dev)
    h_logp = torch.empty(W, dtype=torch.float64).pin_memory()
    h_status = torch.empty(W, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    tstream = torch.cuda.Stream(device=dev)      # the launching stream: kernels and timing events both go here
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    torch.cuda.synchronize()

    def step_dev(i):
        model.loglik_dev(oh, d_theta[i].data_ptr(), W, d_logp.data_ptr(), d_status.data_ptr(), stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------
    for i in range(args.warmup):
        step_dev(i)
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    for k in range(args.steps):
        flush.zero_()                        # L2 flush between timed iterations
        ev[k][0].record()
        step_dev(args.warmup + k)
        ev[k][1].record()
    barrier()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    st_np = d_status.cpu().numpy()
    ok_frac = float((st_np == 0).mean())

    # ---- end-to-end through the host-buffer C ABI (rv_loglik): pinned host memory in, results out ----
    lib = ctx.lib
    import ctypes as C

    def step_e2e(i):
        rc = lib.rv_loglik(ctx.h, model.h, oh.h, C.c_void_p(host_theta[i].data_ptr()), W,
                           C.c_void_p(h_logp.data_ptr()), C.c_void_p(h_status.data_ptr()))
        ctx.check(rc, "rv_loglik")
    step_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(args.warmup + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # keep the GPU under the same load until the clock sampler (nvidia-smi, ~0.3 s per query) has a few samples
    t_load = time.perf_counter()
    while len(sampler.samples) < 4 and time.perf_counter() - t_load < 5.0:
        step_dev(0)
        torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: algorithmic flops / kernel time vs the measured FP64 FMA peak -----------
    ctx.count_work(True)
    ctx.work_counters(reset=True)
    step_dev(0)
    torch.cuda.synchronize()
    S_gpu, T_gpu = ctx.work_counters(reset=True)
    ctx.count_work(False)
    S_eval, T_eval = S_gpu / W, T_gpu / W
    flops_eval = flops_per_eval(S_eval, T_eval)
    kernel_ms = float(ms.mean())             # loglik kernel + finalize (finalize is ~us)
    achieved = flops_eval * W / (kernel_ms * 1e-3) / 1e12
    peak = ctx.fp64_peak_tflops()
    roofline = {"bound": "fp64_fma_pipe", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": 5.33e6, "traffic_note": "dram bytes per launch, profiles/r01j_ncu_loglik_kernel.txt (algorithmic: %d B)" % (W * 92),
                "executed_frac": 0.70,
                "executed_note": "frac uses SURVEY 8(d)'s ALGORITHMIC flop count (rebound's formulation); the kernel executes fewer "
                                 "FP64 instructions per decision (implicit star, coplanar, g-only corrector loop, no divisions), so frac can "
                                 "exceed the pipe's real occupancy: ncu sm__pipe_fp64_cycles_active = 70%% of peak "
                                 "(profiles/r01j_ncu_loglik_kernel.txt)",
                "note": "achieved = SURVEY 8(d) algorithmic flops (S=%.0f force evals, T=%.0f step attempts per eval, "
                        "432*S+1350*T) / CUDA-event kernel time; peak = dependent-free fma.rn.f64 microbenchmark on this GPU "
                        "(rv_fp64_peak; MEASURED_PEAKS.json has no FP64 entry)" % (S_eval, T_eval)}

    # ---- the same evaluations under the two non-default epoch-handling options (never the headline value) ----
    ref_logp = d_logp.clone()
    options = {}
    for key, note in (("monotone_backward", "backward leg visited 0 -> most negative epoch once (state.py:273 order) instead of "
                                            "state.py:91's stored order"),
                      ("dense_output", "one continuous IAS15 integration per leg with natural steps, RVs read from the step's "
                                       "acceleration polynomial instead of a truncated step per epoch")):
        model.set_option(key, 1)
        step_dev(0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); e0.record(); step_dev(0); e1.record(); torch.cuda.synchronize()
        okm = (d_status == 0)
        options[key] = {"value": W / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT + " per GPU", "ms_per_step": e0.elapsed_time(e1),
                        "max_abs_logp_diff_vs_default": float((d_logp[okm] - ref_logp[okm]).abs().max().item()),
                        "note": "model option %s=1: %s; NOT the headline value (the default keeps the reference's step "
                                "sequence)" % (key, note)}
        model.set_option(key, 0)
    step_dev(0); torch.cuda.synchronize()

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:        # a reported baseline: rank 0 at N = 1 only
        cores = os.cpu_count() or 1
        t_probe, _, _, _ = oracle_batch(obs, host_theta[1][:cores * 8].numpy(), cores)      # size the sample: ~15 s of CPU work
        sample = args.cpu_sample or int(min(W, max(cores * 32, 15.0 / max(t_probe, 1e-3) * cores * 8)))
        dt, lo, so, cnt = oracle_batch(obs, host_theta[0][:sample].numpy(), cores)
        # parity spot check on the same vectors
        step_dev(0)
        torch.cuda.synchronize()
        lg = d_logp[:sample].cpu().numpy(); sg = d_status[:sample].cpu().numpy()
        okm = (so == 0) & (sg == 0)
        n1 = 256
        dt1, _, _, _ = oracle_batch(obs, host_theta[2][:n1].numpy(), 1)                 # the same port on ONE core
        cpu_baseline = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "single_core_value": n1 / dt1,
                        "historical_reference": "16.6 evals/s: rebound + emcee under Python 2 on an unknown 2017 CPU, one core, "
                                                "Encounter storms included ((Ex)HD155358.ipynb:181-456; BASELINE.md section 1) -- context only",
                        "sample": "%d walkers of step 0, oracle/rv_oracle.c (gcc -O3 -march=native) OpenMP %d threads, %.1f s" % (sample, cores, dt),
                        "oracle_S_per_eval": cnt[0] / sample, "oracle_T_per_eval": cnt[1] / sample,
                        "parity_max_abs_logp_diff": float(np.abs(lg[okm] - lo[okm]).max()) if okm.any() else None,
                        "parity_status_equal": bool(np.array_equal(so, sg))}

    n_total = W * world * args.steps
    line = {"metric": METRIC, "value": n_total / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic walkers (ensemble start ball) on HD155358.vels epochs",
            "config": {"workload": "HD155358.vels 2-planet log-likelihood (configs[1] shape), %d walkers per GPU per step" % W,
                       "walkers_per_gpu": W, "epochs": 122, "nvars": 10, "integrator": "ias15", "l2": "flushed between steps",
                       "mapping": "lane-per-planet" if args.mapping == 0 else "thread-per-walker", "ok_fraction": ok_frac},
            "e2e": {"value": n_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": W * 10 * 8, "d2h_bytes_per_step": W * 12},
            "gpu_launches": 2 * args.steps, "clocks": sampler.summary(), "roofline": roofline}
    line["non_default_options"] = options
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if args.ess and world == 1:
        line["ess"] = run_ess(ctx, model, oh)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
