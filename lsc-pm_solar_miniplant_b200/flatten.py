"""Scene flattener: pvtrace scene graph -> SoA description (``LscpmSceneDesc`` / ``LscpmBatchDesc``).

The reference hands ``_common_simulation_runner`` a pvtrace ``Scene`` (``miniplant/simulation_runner.py:16-21``)
built by ``miniplant/scene_creator.py:52-274``.  The CUDA library takes plain arrays instead, so this
module walks the graph once: geometry nodes in pre-order with their local->world poses, de-duplicated
materials / components / tabulated spectra, and the light node recognised by the type of its callables
(``LightPosition`` + ``VectorInverter`` = direct beam, ``IsotropicPhotonGenerator`` = sky diffuse,
``PhotonFactory`` = tabulated spectrum; reference ``miniplant/utils.py:39-142``).  Objects are read by
duck typing so real pvtrace objects flatten the same way as the restated ones.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import tally

GEOM = {"Sphere": tally.GEOM_SPHERE, "Box": tally.GEOM_BOX, "Cylinder": tally.GEOM_CYLINDER}
LIGHT_DIRECT, LIGHT_DIFFUSE = 0, 1


class UnsupportedSceneError(ValueError):
    """The scene uses something the flattener cannot describe (unknown geometry, light or callable)."""


def _mro_names(obj) -> list:
    return [c.__name__ for c in type(obj).__mro__]


def _component_kind(comp) -> int:
    names = _mro_names(comp)
    if "Reactor" in names:
        return tally.COMP_REACTOR
    if "Luminophore" in names:
        return tally.COMP_LUMINOPHORE
    if "Absorber" in names or "Scatterer" in names:
        if "Absorber" not in names and float(getattr(comp, "_quantum_yield", 0.0)) > 0:
            raise UnsupportedSceneError("Scatterer components are not on the LSC-PM path")
        return tally.COMP_ABSORBER
    raise UnsupportedSceneError(f"unknown material component {type(comp).__name__}")


@dataclass
class SceneArrays:
    """Flattened scene; field names follow ``LscpmSceneDesc`` (include/lscpm.h)."""

    node_name: list
    node_geom: np.ndarray
    node_parent: np.ndarray
    node_material: np.ndarray
    node_size: np.ndarray
    node_pose: np.ndarray
    mat_refractive_index: np.ndarray
    mat_comp_begin: np.ndarray
    comp_kind: np.ndarray
    comp_coefficient: np.ndarray
    comp_abs_table: np.ndarray
    comp_ems_table: np.ndarray
    comp_quantum_yield: np.ndarray
    table_begin: np.ndarray
    table_x: np.ndarray
    table_y: np.ndarray
    light: Optional["LightArrays"] = None

    @property
    def n_nodes(self) -> int:
        return len(self.node_name)

    def node_components(self, i: int) -> range:
        m = int(self.node_material[i])
        return range(int(self.mat_comp_begin[m]), int(self.mat_comp_begin[m + 1]))

    def bin_names(self) -> list:
        ncomp = [len(self.node_components(i)) for i in range(self.n_nodes)]
        kinds = [[int(self.comp_kind[c]) for c in self.node_components(i)] for i in range(self.n_nodes)]
        return tally.bin_names(self.node_name, self.node_geom, ncomp, kinds)

    def table(self, t: int) -> np.ndarray:
        lo, hi = int(self.table_begin[t]), int(self.table_begin[t + 1])
        return np.stack([self.table_x[lo:hi], self.table_y[lo:hi]], axis=1)


@dataclass
class LightArrays:
    """World-frame description of the scene's light (fields follow ``LscpmBatchDesc``)."""

    kind: int
    mask_half: np.ndarray  # (2,)
    mask_pose: np.ndarray  # (3, 4)
    direction: np.ndarray  # (3,)  DIRECT only
    back_off: float
    tilt_deg: float
    wavelength_nm: float = 555.0  # constant wavelength, used when spectrum is None
    spectrum: Optional[np.ndarray] = None  # (N, 2) knots of the sampling distribution
    wavelength_callable: Optional[Callable] = None  # unrecognised callable: sampled on the host


def flatten_scene(scene, with_light: bool = True) -> SceneArrays:
    names, geoms, parents, materials, sizes, poses = [], [], [], [], [], []
    mat_index, mat_list = {}, []

    def walk(node, parent_world, geom_parent):
        pose = np.asarray(getattr(node, "pose", np.identity(4)), dtype=np.float64)
        world = parent_world @ pose
        geometry = getattr(node, "geometry", None)
        me = geom_parent
        if geometry is not None:
            cls = next((n for n in _mro_names(geometry) if n in GEOM), None)
            if cls is None:
                raise UnsupportedSceneError(f"unsupported geometry {type(geometry).__name__} on node {node.name!r}")
            if cls == "Box":
                size = tuple(float(s) for s in geometry.size)
            elif cls == "Cylinder":
                size = (float(geometry.length), float(geometry.radius), 0.0)
            else:
                size = (float(geometry.radius), 0.0, 0.0)
            material = geometry.material
            if id(material) not in mat_index:
                mat_index[id(material)] = len(mat_list)
                mat_list.append(material)
            me = len(names)
            names.append(str(node.name))
            geoms.append(GEOM[cls])
            parents.append(geom_parent)
            materials.append(mat_index[id(material)])
            sizes.append(size)
            poses.append(world)
        for child in node.children:
            walk(child, world, me)

    walk(scene.root, np.identity(4), -1)
    if not names or getattr(scene.root, "geometry", None) is None:
        raise UnsupportedSceneError("the scene root must carry the world geometry")

    # materials -> components -> tables, de-duplicated by identity of the source arrays
    comp_kind, comp_coef, comp_abs, comp_ems, comp_qy = [], [], [], [], []
    mat_comp_begin = [0]
    tables, table_index = [], {}

    def table_id(arr) -> int:
        arr = np.asarray(arr, dtype=np.float64)
        key = arr.tobytes()
        if key not in table_index:
            if arr.ndim != 2 or arr.shape[1] != 2 or len(arr) < 2:
                raise UnsupportedSceneError("tabulated spectra must be (N>=2, 2) arrays")
            if np.any(np.diff(arr[:, 0]) <= 0):
                raise UnsupportedSceneError("spectrum abscissae must be strictly increasing")
            if np.any(arr[:, 1] < 0):
                raise UnsupportedSceneError("spectrum ordinates must be non-negative")
            table_index[key] = len(tables)
            tables.append(arr)
        return table_index[key]

    for material in mat_list:
        for comp in material.components:
            kind = _component_kind(comp)
            coef = comp._coefficient
            comp_kind.append(kind)
            if np.isscalar(coef):
                comp_coef.append(float(coef))
                comp_abs.append(-1)
            else:
                comp_coef.append(0.0)
                comp_abs.append(table_id(coef))
            if kind == tally.COMP_LUMINOPHORE:
                dist = comp._ems_dist
                comp_ems.append(table_id(np.stack([np.asarray(dist._x, float), np.asarray(dist._y, float)], axis=1)))
                comp_qy.append(float(comp._quantum_yield))
            else:
                comp_ems.append(-1)
                comp_qy.append(0.0)
        mat_comp_begin.append(len(comp_kind))

    table_begin = np.cumsum([0] + [len(t) for t in tables]).astype(np.int32)
    flat = np.concatenate(tables, axis=0) if tables else np.zeros((0, 2))
    arrays = SceneArrays(
        node_name=names,
        node_geom=np.asarray(geoms, dtype=np.int32),
        node_parent=np.asarray(parents, dtype=np.int32),
        node_material=np.asarray(materials, dtype=np.int32),
        node_size=np.ascontiguousarray(np.asarray(sizes, dtype=np.float64).reshape(-1, 3)),
        node_pose=np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 16)),
        mat_refractive_index=np.asarray([float(m.refractive_index) for m in mat_list], dtype=np.float64),
        mat_comp_begin=np.asarray(mat_comp_begin, dtype=np.int32),
        comp_kind=np.asarray(comp_kind, dtype=np.int32),
        comp_coefficient=np.asarray(comp_coef, dtype=np.float64),
        comp_abs_table=np.asarray(comp_abs, dtype=np.int32),
        comp_ems_table=np.asarray(comp_ems, dtype=np.int32),
        comp_quantum_yield=np.asarray(comp_qy, dtype=np.float64),
        table_begin=table_begin,
        table_x=np.ascontiguousarray(flat[:, 0]),
        table_y=np.ascontiguousarray(flat[:, 1]),
    )
    if with_light:
        arrays.light = flatten_light(scene)
    return arrays


# --------------------------------------------------------------------------------------
# light
# --------------------------------------------------------------------------------------


def _world_pose(node) -> np.ndarray:
    m = np.identity(4)
    for n in node.path:
        m = m @ np.asarray(getattr(n, "pose", np.identity(4)), dtype=np.float64)
    return m


def _mask_matrix(position) -> np.ndarray:
    """Mask plane -> light-node frame, 4x4, of a ``LightPosition``.

    The reference's class (``miniplant/utils.py:87-99``) exposes only ``tilt_angle`` and ``__call__``; the matrix it
    builds per photon is ``inv(R_y(-tilt))`` (``utils.py:94``), i.e. a rotation by +tilt about y.  Derived here from
    ``tilt_angle`` so the reference's unmodified objects flatten; a ``matrix()`` method is not required."""
    try:
        t = np.radians(float(position.tilt_angle))
    except (AttributeError, TypeError, ValueError) as exc:
        raise UnsupportedSceneError(f"light position {type(position).__name__} exposes no numeric tilt_angle") from exc
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0.0, s, 0.0], [0.0, 1.0, 0.0, 0.0], [-s, 0.0, c, 0.0], [0.0, 0.0, 0.0, 1.0]])


def _wavelength_source(fn, light: LightArrays) -> None:
    """Recognise the wavelength callable: a ``PhotonFactory`` over a ``Distribution`` (tabulated
    spectrum), or a callable returning a constant; anything else is sampled on the host."""
    spectrum = getattr(fn, "spectrum", None)
    if spectrum is not None and hasattr(spectrum, "_x") and hasattr(spectrum, "_y"):
        x = np.asarray(spectrum._x, dtype=np.float64)
        y = np.asarray(spectrum._y, dtype=np.float64)
        if getattr(spectrum, "hist", False):
            raise UnsupportedSceneError("histogram spectra are not on the LSC-PM path")
        light.spectrum = np.stack([x, y], axis=1)
        return
    # Only provably constant sources go to the device: plain numbers and functions whose bytecode is nothing
    # but "return <literal or variable>" (the reference's default ``lambda: 555``, simulation_runner.py:87,111).  Any
    # other callable may draw from a generator this module cannot see, so it is sampled on the host by the
    # scene's own ``emit`` — never called here (no side effects, no guessing from a few samples).
    constant = _constant_value(fn)
    if constant is not None:
        light.wavelength_nm = constant
    else:
        light.wavelength_callable = fn


# Opcodes of a function body that is exactly ``return <literal | closure variable | global>``: no call, no
# attribute access, no arithmetic — evaluating it once has no side effects and consumes no random numbers.
_CONSTANT_OPS = frozenset({"RESUME", "LOAD_CONST", "RETURN_VALUE", "RETURN_CONST", "NOP", "CACHE", "COPY_FREE_VARS",
                           "LOAD_DEREF", "LOAD_GLOBAL", "PUSH_NULL"})


def _constant_value(fn) -> Optional[float]:
    if isinstance(fn, (int, float, np.integer, np.floating)) and not isinstance(fn, bool):
        return float(fn)
    if "ConstantWavelength" in _mro_names(fn):  # workloads.ConstantWavelength declares itself constant
        return float(fn.nm)
    code = getattr(fn, "__code__", None)
    if code is None or code.co_argcount or code.co_kwonlyargcount:
        return None
    import dis

    if any(ins.opname not in _CONSTANT_OPS for ins in dis.get_instructions(code)):
        return None
    value = fn()
    if isinstance(value, (int, float, np.integer, np.floating)) and not isinstance(value, bool):
        return float(value)
    return None


def flatten_light(scene) -> LightArrays:
    """Light node -> ``LightArrays``.  Anything that is not one of the reference's two light constructions raises
    ``UnsupportedSceneError`` (also for objects that miss an expected attribute), which ``backend.simulate_scene``
    answers by emitting on the host through the scene's own ``emit``."""
    try:
        return _flatten_light(scene)
    except (AttributeError, TypeError, IndexError) as exc:
        raise UnsupportedSceneError(f"unrecognised light: {exc}") from exc


def _flatten_light(scene) -> LightArrays:
    lights = scene.light_nodes
    if len(lights) != 1:
        raise UnsupportedSceneError(f"the LSC-PM path expects exactly one light node, found {len(lights)}")
    node = lights[0]
    source = node.light
    world = _world_pose(node)
    position = getattr(source, "position", None)
    direction = getattr(source, "direction", None)
    joint = getattr(source, "position_direction", None)

    if joint is not None and "IsotropicPhotonGenerator" in _mro_names(joint):
        base = joint.base_position_generator
        mask = world @ _mask_matrix(base)
        light = LightArrays(kind=LIGHT_DIFFUSE, mask_half=np.array([0.47 / 2, 0.47 / 2]), mask_pose=mask[:3, :],
                            direction=np.zeros(3), back_off=1.0, tilt_deg=float(joint.tilt_angle))
        if not np.allclose(world[:3, :3], np.identity(3)):
            raise UnsupportedSceneError("a rotated diffuse light node is not on the LSC-PM path")
    elif position is not None and "LightPosition" in _mro_names(position) \
            and direction is not None and "VectorInverter" in _mro_names(direction):
        d_local = -np.asarray(direction.vector, dtype=np.float64)
        d_world = world[:3, :3] @ d_local
        norm = float(np.linalg.norm(d_world))
        if not np.isclose(norm, 1.0, atol=1e-9):
            raise UnsupportedSceneError("the beam direction must be a unit vector")
        mask = world @ _mask_matrix(position)
        # The light node is translated up-beam (scene_creator.py:304): express that as a back-off
        # along the ray so that the anchor points stay on the plate face.
        shift = world[:3, 3]
        back_off = -float(np.dot(shift, d_world))
        if np.allclose(shift, -back_off * d_world, atol=1e-12) and back_off >= 0:
            mask = mask.copy()
            mask[:3, 3] -= shift
        else:
            back_off = 0.0
        light = LightArrays(kind=LIGHT_DIRECT, mask_half=np.array([0.47 / 2, 0.47 / 2]), mask_pose=mask[:3, :],
                            direction=d_world / norm, back_off=back_off, tilt_deg=float(position.tilt_angle))
    else:
        raise UnsupportedSceneError(
            "unrecognised light: the CUDA path samples LightPosition+VectorInverter (direct) and "
            "IsotropicPhotonGenerator (diffuse) lights on the device")
    _wavelength_source(source.wavelength, light)
    return light
