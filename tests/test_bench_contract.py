"""The driver depends on bench.py's one-line JSON contract.  The reference arm (--impl reference) needs no GPU, so its
line is checked here on the CPU; the GPU arm's line is checked in test_bench_contract_gpu (marker gpu)."""
import json
import os
import subprocess
import sys

import pytest

import rvtest as T

BENCH = os.path.join(T.ROOT, "bench.py")
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e"}


def _one_line(args, timeout):
    r = subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _one_line(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert COMMON <= set(d) and d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["value"] > 10 and "workload" in d["config"] and d["vs_baseline"] is None and d["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _one_line(["--steps", "3", "--warmup", "3", "--walkers", "8192", "--cpu-sample", "256", "--ess-rows", "60",
                   "--ess-budget", "4,3,4"], 900)
    assert COMMON <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["gpu_launches"] == 15 and d["value"] > 1e5          # 3 steps x (3 item-order kernels + likelihood + finalize)
    assert d["e2e"]["h2d_bytes_per_step"] == 8192 * 80 and d["e2e"]["d2h_bytes_per_step"] == 8192 * 12 and d["e2e"]["value"] > 1e5
    rf = d["roofline"]
    assert rf["bound"] == "fp64_fma_pipe" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["unit"] == "TFLOP/s"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["parity_status_equal"] and cb["parity_max_abs_logp_diff"] < 1e-6
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # side blocks: ESS/s of the three samplers from the equilibrated ensemble, the variational kernel, profile-sourced figures
    for k in ("stretch", "mh", "smala"):
        assert d["ess"][k]["ess_per_s"] > 0 and d["ess"][k]["recorded_rows"] >= 50 and 0.05 < d["ess"][k]["accept_rate"] < 0.95
    assert d["var"]["value"] > 1e3 and 0 < d["var"]["roofline"]["frac"] < 1.2
    assert rf["from_profile"]["file"] == "profiles/roofline_inputs.json" and "source" in rf["from_profile"]["loglik_kernel"]
