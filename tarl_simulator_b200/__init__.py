"""tarl_simulator_b200 — B200-native (sm_100a) implementation of TARL-simulator's data-parallel hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute of the path runs in
hand-written CUDA kernels reached through the C ABI in include/tarl_b200.h (libtarl_b200.so, loaded with ctypes).
There is no CPU fallback: using the compute classes without the built library or without a CUDA device raises.

The reference's import paths (`src.simulation_core_model`, `src.direction_mpnn`, …) are mirrored by the thin `src/`
package at the repo root, which re-exports the classes defined here.
"""
from .feature_helpers import AgentFeatureHelpers, FeatureHelpers, ObservationFeatureHelpers  # noqa: F401
from .data import Data  # noqa: F401

__all__ = ["FeatureHelpers", "AgentFeatureHelpers", "ObservationFeatureHelpers", "Data"]
