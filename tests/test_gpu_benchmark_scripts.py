"""The reference's stand-alone workloads on the GPU backend: bench_scripts/mcmc_benchmarks.py {mh, smala, emcee}
(Py3 equivalents of mcmc_benchmark_mh.py:32-60, mcmc_benchmark_smala.py:32-54, mcmc_benchmark_emcee.py:33-55 -- the
originals cannot run even on the reference) and examples/usage_example.py (the "(Ex)Full Test + Usage Example"
notebook experiment), at small iteration counts: they must run to completion through the State / Observations / mcmc API,
print finite posterior means and AC times, and the usage example's cross-sampler KS distances must sit in the notebook's
range ((Ex)Full Test + Usage Example.ipynb:710-718: D = 0.014-0.050 between samplers)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import rvtest as T

pytestmark = pytest.mark.gpu
SCRIPT = os.path.join(T.ROOT, "bench_scripts", "mcmc_benchmarks.py")


def _run(args, timeout=600):
    r = subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=timeout, cwd=T.ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    return r.stdout


def _floats(line):
    return [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line.split(":", 1)[1])]


def _vector_after(out, label):
    """numpy prints a 10-vector over several lines: collect everything between `label` and the closing bracket."""
    tail = out[out.index(label) + len(label):]
    return _floats("x:" + tail[:tail.index("]")])


@pytest.mark.parametrize("which,niter,nvars,fused", [
    ("mh", 300, 10, True), ("smala", 60, 10, True), ("emcee", 32 * 40, 10, True),
    ("smala", 12, 10, False), ("emcee", 32 * 6, 10, False),
])
def test_benchmark_script_workloads_run_on_gpu(which, niter, nvars, fused):
    out = _run([SCRIPT, which, "--niter", str(niter)] + (["--fused"] if fused else []))
    mean = _vector_after(out, "mean:")
    true = _vector_after(out, "true:")
    assert len(mean) == nvars and len(true) == nvars and np.all(np.isfinite(mean))
    # short chains started at the truth stay near it (every parameter within 20 % of its scale or 0.2 absolute)
    assert np.all(np.abs(np.array(mean) - np.array(true)) < 0.2 * np.maximum(np.abs(true), 1.0))
    acs = [int(l.rsplit(":", 1)[1]) for l in out.splitlines() if l.startswith("AC time")]
    assert len(acs) == nvars and all(a >= 1 for a in acs)
    assert "iterations in" in out


def test_mh_benchmark_script_step_force_loop():
    # mcmc_benchmark_mh.py:58-61 records a row only after an acceptance (mh.step_force); with the reference's scales the
    # acceptance is low, so only a handful of rows are asked for
    out = _run([SCRIPT, "mh", "--niter", "4"], timeout=900)
    assert "Acceptance rate" in out and "AC time" in out


def test_usage_example_cross_sampler_agreement():
    out = _run([os.path.join(T.ROOT, "examples", "usage_example.py"), "--niter", "1200", "--chains", "128"], timeout=900)
    rows = {l.split("vs")[1].strip().rsplit("[", 1)[0].strip(): _floats("x:" + l.rsplit("[", 1)[1])
            for l in out.splitlines() if l.startswith("KS distance")}
    for name in ("emcee fused", "smala fused", "alsmala fused"):
        assert max(rows[name]) < 0.06, (name, rows[name])            # notebook: 0.014-0.050
    ac = {l.split("  ")[0].strip(): l for l in out.splitlines() if "AC times" in l}
    smala_ac = _floats("x:" + ac["smala fused"].split("AC times")[1].split("efficacy")[0])
    assert max(smala_ac) <= 3.0                                        # notebook: SMALA AC 1/1/1
    assert "logp(true)" in out
