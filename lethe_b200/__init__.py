"""lethe_b200 — B200-native DEM time-step hot path of lethe-particles.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of
include/lethe_dem.h), `abi` (ctypes binding), `prm` (the reference's `.prm`
parameter interface) and `solver` (host-side mirror of DEMSolver).
"""
from . import abi, prm, solver  # noqa: F401
from .abi import Config, DEMError, Engine, load_engine  # noqa: F401
from .prm import DEMParameters, load_prm, parameters_from_prm  # noqa: F401
from .solver import DEMSolver  # noqa: F401
