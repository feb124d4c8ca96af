"""The drop-in boundary is a C ABI: include/rvgpu.h must compile as C99, and a plain-C client (no Python, no torch in the
process) must be able to drive the library."""
import os
import subprocess

import pytest

import rvtest as T

INC = os.path.join(T.ROOT, "include")
SRC = os.path.join(T.ROOT, "tests", "c_abi", "abi_smoke.c")
LIBDIR = os.path.join(T.ROOT, "rvel_mcmc_b200")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def test_header_is_valid_c99():
    r = subprocess.run([GCC, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", INC, "-x", "c",
                        os.path.join(INC, "rvgpu.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_client_compiles_and_links():
    out = os.path.join(T.ROOT, "tests", "c_abi", "_build")
    os.makedirs(out, exist_ok=True)
    r = subprocess.run([GCC, "-std=c99", "-Wall", "-O1", "-I", INC, SRC, "-o", os.path.join(out, "abi_smoke"),
                        "-L", LIBDIR, "-lrvgpu", "-Wl,-rpath," + LIBDIR, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_c_client_runs_on_gpu():
    test_c_client_compiles_and_links()
    exe = os.path.join(T.ROOT, "tests", "c_abi", "_build", "abi_smoke")
    r = subprocess.run([exe, os.path.join(T.GOLDEN, "HD155358.vels")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    f = r.stdout.split()
    assert f[1] == "0" and f[5] == "0"
    assert abs(float(f[3]) - T.KAT2_LOGP) < 5e-11 and abs(float(f[7]) - T.KAT2_LOGP) < 5e-11      # KAT-2 through plain C
    assert abs(float(f[9]) / 307.60027893 - 1) < 1e-8                                          # SURVEY B.9 dlogp/da0
