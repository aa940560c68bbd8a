"""Host-side compatibility layer: the third-party names the reference's hot path imports
(``pvtrace``, ``anytree``, ``pvlib.irradiance``) restated as plain scene-description objects.

``install()`` registers them under their upstream module names **only when the real package is
not importable**, so reference-style code (``from pvtrace import Node, Box, ...``) runs unchanged.
"""
from __future__ import annotations

import importlib
import sys
import types

from . import anytree_lite, pvlib_lite, pvtrace_api

_PVTRACE_NAMES = (
    "Node", "Box", "Sphere", "Cylinder", "Material", "Surface", "Component", "Scatterer", "Absorber",
    "Luminophore", "Reactor", "Light", "Scene", "Distribution", "Ray", "Event", "isotropic",
    "rectangular_mask", "MeshcatRenderer",
)


def _importable(name: str) -> bool:
    if name in sys.modules:
        return not getattr(sys.modules[name], "__lscpm_shim__", False)
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__lscpm_shim__ = True
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install(force: bool = False) -> dict:
    """Register shim modules for ``pvtrace``, ``anytree`` and ``pvlib`` where the real ones are missing.

    Returns ``{module_name: "real" | "shim"}`` so callers and tests can see what is in use."""
    status = {}
    if force or not _importable("anytree"):
        _module("anytree", NodeMixin=anytree_lite.NodeMixin, Node=anytree_lite.Node,
                LevelOrderIter=anytree_lite.LevelOrderIter, PreOrderIter=anytree_lite.PreOrderIter,
                LoopError=anytree_lite.LoopError)
        status["anytree"] = "shim"
    else:
        status["anytree"] = "real"

    if force or not _importable("pvtrace"):
        from .. import photon_tracer as _pt

        root = _module("pvtrace", **{n: getattr(pvtrace_api, n) for n in _PVTRACE_NAMES})
        root.photon_tracer = _pt
        root.__path__ = []  # behave as a package so sub-module imports resolve through sys.modules
        geometry = _module("pvtrace.geometry")
        geometry.__path__ = []
        tf = _module("pvtrace.geometry.transformations",
                     rotation_matrix=pvtrace_api.rotation_matrix,
                     translation_matrix=pvtrace_api.translation_matrix)
        geometry.transformations = tf
        material = _module("pvtrace.material")
        material.__path__ = []
        mutils = _module("pvtrace.material.utils", spherical_to_cart=pvtrace_api.spherical_to_cart,
                         isotropic=pvtrace_api.isotropic)
        material.utils = mutils
        algorithm = _module("pvtrace.algorithm")
        algorithm.__path__ = []
        sys.modules["pvtrace.algorithm.photon_tracer"] = _pt
        algorithm.photon_tracer = _pt
        root.geometry, root.material, root.algorithm = geometry, material, algorithm
        status["pvtrace"] = "shim"
    else:
        status["pvtrace"] = "real"

    if force or not _importable("pvlib"):
        root = _module("pvlib", irradiance=pvlib_lite.irradiance)
        root.__path__ = []
        loc = _module("pvlib.location", Location=pvlib_lite.Location)
        root.location = loc
        sys.modules["pvlib.irradiance"] = _module("pvlib.irradiance",
                                                  aoi_projection=pvlib_lite.irradiance.aoi_projection,
                                                  aoi=pvlib_lite.irradiance.aoi)
        root.irradiance = sys.modules["pvlib.irradiance"]
        status["pvlib"] = "shim"
    else:
        status["pvlib"] = "real"
    return status


def pvtrace_namespace():
    """The object API the ``miniplant`` mirror builds scenes with: real ``pvtrace`` when it is
    importable, otherwise the restated scene-description classes."""
    if _importable("pvtrace"):
        import pvtrace  # type: ignore
        from pvtrace.geometry.transformations import rotation_matrix  # type: ignore
        from pvtrace.material.utils import spherical_to_cart  # type: ignore

        ns = types.SimpleNamespace(**{n: getattr(pvtrace, n) for n in _PVTRACE_NAMES if hasattr(pvtrace, n)})
        ns.rotation_matrix = rotation_matrix
        ns.spherical_to_cart = spherical_to_cart
        ns.is_shim = False
        return ns
    ns = types.SimpleNamespace(**{n: getattr(pvtrace_api, n) for n in _PVTRACE_NAMES})
    ns.rotation_matrix = pvtrace_api.rotation_matrix
    ns.spherical_to_cart = pvtrace_api.spherical_to_cart
    ns.is_shim = True
    return ns
