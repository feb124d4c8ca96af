"""rvel_mcmc_b200 -- B200-native (sm_100a) hot path of rvel-mcmc.

The reference's Python API (state.State, observations.*, mcmc.*, driver.*) is kept; the N-body
RV log-likelihood underneath runs in hand-written CUDA kernels reached through the C ABI of
``librvgpu.so`` (include/rvgpu.h).  There is no CPU fallback: importing this package works without a
GPU (host logic only), but any likelihood evaluation raises if the library or a B200 is missing.
"""
from ._abi import Encounter, RvGpuError, lib_path  # noqa: F401
from . import observations, state, mcmc, driver  # noqa: F401
from .state import State  # noqa: F401

__all__ = ["State", "observations", "state", "mcmc", "driver", "Encounter", "RvGpuError"]
