"""CPU-side check of the drop-in boundary: librvgpu.so loads and exports every entry point that
include/rvgpu.h declares; without a GPU a call fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import rvtest as T
from rvel_mcmc_b200 import _abi


def _declared():
    src = open(os.path.join(T.ROOT, "include", "rvgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_abi.lib_path()), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(_abi.lib_path())
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    # the Python binding covers the same set
    assert set(_abi.exported_symbols()) == set(names)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_abi.RvGpuError):
        _abi.Context(0)


def test_product_does_not_link_the_oracle():
    # the shipped library and package must not reference oracle/ or the host mirror
    pkg = os.path.join(T.ROOT, "rvel_mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "librvoracle" not in txt and "hostmirror" not in txt and "rv_oracle" not in txt, f


def test_default_context_fails_loudly_without_a_gpu_and_does_not_deadlock():
    """default_context() takes the module lock and then constructs a Context, which loads the library under the same lock."""
    import subprocess
    import sys
    import rvtest as T
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from rvel_mcmc_b200 import _abi\n"
            "try:\n    _abi.default_context(); print('CTX')\n"
            "except _abi.RvGpuError as e:\n    print('RvGpuError')\n") % T.ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() in ("RvGpuError", "CTX")          # CTX on a GPU box, RvGpuError here; never a hang
