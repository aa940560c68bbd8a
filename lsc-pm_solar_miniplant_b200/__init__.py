"""lsc-pm_solar_miniplant_b200 — B200-native Monte-Carlo photon tracer for the LSC-PM hot path.

The directory name follows the project naming rule and is not a valid Python identifier; import it
through the root-level alias ``import lscpm_b200`` (``lscpm_b200.py`` loads this package under that
name), or with ``importlib.import_module("lsc-pm_solar_miniplant_b200")``.

Layout (only what the hot path needs):

``csrc/``          sm_100a CUDA kernels + the C-ABI of ``include/lscpm.h`` (built to ``liblscpm.so``)
``_native.py``     ctypes binding of the C-ABI; fails loudly when the library is missing
``backend.py``     device-side plumbing (torch tensors/streams), scene handles, launches
``flatten.py``     scene graph -> SoA description (``LscpmSceneDesc``)
``tally.py``       naming contract of the per-node tally bins
``distributed.py`` batch partitioning across ranks + NCCL all-reduce of the integer tallies
``compat/``        the pvtrace / anytree / pvlib names the reference imports, as scene-description objects
``miniplant/``     drop-in mirror of the reference's ``miniplant`` package for this path
"""
from __future__ import annotations

import sys as _sys

__version__ = "0.1.0"


def install(force: bool = False) -> dict:
    """Make reference-style imports work: registers ``miniplant`` (this package's mirror) and, where
    the real packages are missing, ``pvtrace`` / ``anytree`` / ``pvlib`` shims in ``sys.modules``."""
    from . import compat, miniplant

    status = compat.install(force=force)
    if force or "miniplant" not in _sys.modules:
        _sys.modules["miniplant"] = miniplant
        for sub in ("utils", "scene_creator", "simulation_runner", "solar_data", "locations",
                    "full_simulation", "angle_optimization_simulation"):
            try:
                mod = __import__(f"{__name__}.miniplant.{sub}", fromlist=[sub])
            except ImportError:
                continue
            _sys.modules[f"miniplant.{sub}"] = mod
        status["miniplant"] = "mirror"
    return status
