"""rfsurfhmc_b200 — B200-native (sm_100a CUDA) drop-in for the hot path of nqdu/RfSurfHmc.

Layout (mirrors the reference so that `from model.lib import libsurf`-style code ports 1:1):
  rfsurfhmc_b200.model.lib.libsurf / librf   pybind11-signature drop-ins (src/SWD/main.cpp, src/RF/main.cpp)
  rfsurfhmc_b200.model.model_rf / model_surf / model_rf_swd_vs_thk   (model/*.py)
  rfsurfhmc_b200.pyhmc.hmc / hmcda                                   (pyhmc/*.py)
  rfsurfhmc_b200.batched    batched device API (torch tensors in, torch tensors out)
  rfsurfhmc_b200.csrc       hand-written CUDA kernels + the C ABI (include/rfsurfhmc.h)

There is no CPU fallback: importing the numerical entry points without the compiled
librfsurf_b200.so, or calling them without a CUDA device, raises.
"""
from ._lib import RfsError, load_library, Context  # noqa: F401

__version__ = "0.1.0"
