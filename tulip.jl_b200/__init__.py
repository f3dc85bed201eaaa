"""tulip.jl_b200 -- B200-native KKT linear-algebra backend for Tulip.jl's interior-point method.

Scope: the per-IPM-iteration hot path behind ``KKT.update!`` / ``KKT.solve!``
-- assemble A*Theta*A' (K1) or the augmented matrix (K2), supernodal numeric factorisation,
triangular solves -- as hand-written CUDA for sm_100a behind a C ABI (include/tlpb200.h).

The directory name contains a dot, so it is loaded through ``tlpb200_loader`` (repo root), which
registers it under the importable name ``tulip_jl_b200``.
"""
from . import _lib, hsd, hsd_device, ipmdata, kkt, lpgen, mpc  # noqa: F401
from .hsd_device import DeviceHSD  # noqa: F401
from .kkt import (K1, K2, Backend, B200KKTSolver, DefaultKKTSystem, DimensionMismatch,  # noqa: F401
                  OutOfMemoryError, PosDefException, TlpB200Error, arithmetic, backend, linear_system,
                  setup, solve_, update_)

__all__ = ["K1", "K2", "Backend", "B200KKTSolver", "DefaultKKTSystem", "setup", "update_", "solve_",
           "arithmetic", "backend", "linear_system", "PosDefException", "DimensionMismatch",
           "OutOfMemoryError", "TlpB200Error", "lpgen", "kkt", "hsd", "hsd_device", "DeviceHSD", "mpc", "ipmdata"]
