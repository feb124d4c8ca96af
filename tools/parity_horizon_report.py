#!/usr/bin/env python3
"""Runs the 10^4-step decision-identity comparisons of tests/parity_horizon.py (device samplers through the C ABI against
the CPU oracle, same Philox streams) and prints one JSON object per case with the first diverging step (null = identical
over the whole run).  The committed copy of its output is profiles/r02_parity_horizon.jsonl."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_horizon as PH
import rvtest as T
from rvel_mcmc_b200 import _abi

ctx = _abi.Context(0)
cases = [
    lambda: PH.mh_horizon(ctx, "hd155358", 16, 10000, T.HD_SCALE_VEC, 0.3, seed=99,
                          theta0=T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 16, 4)),
    lambda: PH.stretch_horizon(ctx, "hd155358", 32, 10000, seed=5),
    lambda: PH.stretch_horizon(ctx, "small", 64, 10000, seed=77, width=1.0, ball_seed=1),
    lambda: PH.mh_horizon(ctx, "small", 16, 10000, [3e-4, 0.01, 0.01], 5.0, seed=2024),
]
for c in cases:
    print(json.dumps(c()), flush=True)
