#!/usr/bin/env python3
"""Record of the roofline denominator: the FP64 FMA-pipe peak of this GPU as rv_fp64_peak measures it (MEASURED_PEAKS.json has
no FP64 entry), with the clocks it was measured at and the theoretical figure beside it.  Prints one JSON object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rvel_mcmc_b200 import _abi

ctx = _abi.Context(0)
vals = [ctx.fp64_peak_tflops() for _ in range(5)]
info = ctx.device_info()
q = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                    "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print(json.dumps({"tflops_best": max(vals), "tflops_runs": vals, "sm_count": info["sm_count"],
                  "theoretical_tflops_at_max_clock": info["sm_count"] * 64 * 2 * info["clock_khz"] * 1e3 / 1e12,
                  "clock_khz_attr": info["clock_khz"], "nvidia_smi": q,
                  "method": "rv_fp64_peak (rv_kernels.cu fp64_peak_kernel): sm_count x 8 CTAs x 256 threads, 8 independent "
                            "fma.rn.f64 chains per thread x 4096 x 8 iterations, CUDA events on the context stream, best of 5 "
                            "after one warm-up; flops = 2 per FMA"}))
