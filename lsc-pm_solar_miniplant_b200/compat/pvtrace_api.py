"""Scene-description half of the pvtrace object API (host logic only).

The reference builds its scene with pvtrace objects (``miniplant/scene_creator.py:8-21``,
``miniplant/utils.py:8-11``) and hands the resulting ``Scene`` to the hot path
(``miniplant/simulation_runner.py:16-80``).  pvtrace (fork ``cambiegroup/pvtrace@22c9076``,
reference ``requirements.txt:43``) is a third-party dependency that is not vendored in the
reference and not installable here, so this module restates the *scene-description* classes
with the constructor signatures and attributes the reference's call sites use.  They are plain
data holders: **no ray-tracing arithmetic lives here**.  The tracing itself is done by the CUDA
library behind ``include/lscpm.h`` (product) or by ``oracle/`` (test infrastructure only).

Behavioural notes each class follows (pvtrace 2.1 public behaviour):

* ``Node.translate``/``Node.rotate`` pre-multiply the pose, i.e. act in the parent frame; the
  reference's tilt transform only puts the plate's top face through the world origin under that
  convention (``miniplant/scene_creator.py:264-272``).
* ``Distribution`` builds its CDF as ``numpy.cumsum(y) / max`` over the knots (pvtrace's ``cdf = np.cumsum(y);
  pdf = cdf / max(cdf)`` idiom, the same one ``Material.component`` uses) — knot spacing does not enter, which
  matters on the non-uniform 27-point solar grid; both ``lookup`` and ``sample`` are ``numpy.interp`` on the
  knots; a negative ordinate raises ``ValueError`` (relied on by ``miniplant/solar_data.py:93-104``).  The choice
  between cumsum and a cumulative trapezoid is pinned by the reference's committed results.
* ``Light.emit`` calls the wavelength / position / direction callables once per ray
  (``miniplant/scene_creator.py:295-303``).
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, replace
from typing import Callable, Iterator, Optional, Sequence

import numpy as np

from .anytree_lite import NodeMixin, LevelOrderIter

# --------------------------------------------------------------------------------------
# transformations (pvtrace.geometry.transformations subset)
# --------------------------------------------------------------------------------------


def translation_matrix(direction) -> np.ndarray:
    m = np.identity(4)
    m[:3, 3] = np.asarray(direction, dtype=float)[:3]
    return m


def rotation_matrix(angle, direction, point=None) -> np.ndarray:
    """4x4 homogeneous rotation by ``angle`` (radians) about ``direction`` (Rodrigues)."""
    axis = np.asarray(direction, dtype=float)[:3]
    axis = axis / math.sqrt(float(np.dot(axis, axis)))
    s, c = math.sin(angle), math.cos(angle)
    r = np.diag([c, c, c]) + np.outer(axis, axis) * (1.0 - c)
    ax, ay, az = axis * s
    r += np.array([[0.0, -az, ay], [az, 0.0, -ax], [-ay, ax, 0.0]])
    m = np.identity(4)
    m[:3, :3] = r
    if point is not None:
        p = np.asarray(point, dtype=float)[:3]
        m[:3, 3] = p - r @ p
    return m


def spherical_to_cart(theta, phi, r=1.0) -> np.ndarray:
    """(theta = polar angle from +z, phi = azimuth from +x) -> cartesian vector."""
    return np.array(
        [
            r * np.sin(theta) * np.cos(phi),
            r * np.sin(theta) * np.sin(phi),
            r * np.cos(theta),
        ]
    )


# --------------------------------------------------------------------------------------
# sampling helpers
# --------------------------------------------------------------------------------------


def isotropic() -> tuple:
    """Uniform direction on the unit sphere: phi=U(0,2pi), mu=U(-1,1)."""
    phi = np.random.uniform(0.0, 2.0 * np.pi)
    mu = np.random.uniform(-1.0, 1.0)
    theta = np.arccos(mu)
    return (
        float(np.sin(theta) * np.cos(phi)),
        float(np.sin(theta) * np.sin(phi)),
        float(np.cos(theta)),
    )


def rectangular_mask(l, w) -> tuple:
    """Uniform point on the z=0 rectangle |x|<=l, |y|<=w (half-widths)."""
    return (np.random.uniform(-l, l), np.random.uniform(-w, w), 0.0)


class Distribution:
    """Tabulated 1-D distribution with a cumulative-sum CDF; see module docstring."""

    def __init__(self, x=None, y=None, hist: bool = False):
        if x is None or y is None:
            raise ValueError("Distribution needs both x and y")
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        if np.any(y < 0):
            raise ValueError("Distribution ordinates must be non-negative.")
        self.hist = hist
        self._x = x
        self._y = y
        if hist:
            if len(x) != len(y) + 1:
                raise ValueError("Histogram distributions need len(x) == len(y) + 1")
            cdf = np.hstack([0.0, np.cumsum(y * np.diff(x))])
        else:
            cdf = np.cumsum(y)
        peak = cdf.max() if len(cdf) else 0.0
        self._cdf = cdf / peak if peak > 0 else cdf

    def __call__(self, x):
        if self.hist:
            idx = np.clip(np.searchsorted(self._x, x, side="right") - 1, 0, len(self._y) - 1)
            return self._y[idx]
        return np.interp(x, self._x, self._y)

    def lookup(self, x):
        """CDF value at abscissa ``x``."""
        return np.interp(x, self._x, self._cdf)

    def sample(self, u):
        """Inverse-CDF: abscissa at cumulative probability ``u``."""
        return np.interp(u, self._cdf, self._x)


class LazyDistribution(Distribution):
    """A non-histogram ``Distribution`` over already validated knots whose CDF is built on first use.  The solar stage
    creates two per time point (tens of thousands per year); the drivers only read ``_x`` / ``_y`` from them."""

    def __init__(self, x, y):   # noqa: super().__init__ deliberately not called
        self.hist = False
        self._x = x
        self._y = y

    @property
    def _cdf(self):
        cdf = self.__dict__.get("_cdf_cache")
        if cdf is None:
            cdf = np.cumsum(self._y)
            peak = cdf.max() if len(cdf) else 0.0
            cdf = cdf / peak if peak > 0 else cdf
            self.__dict__["_cdf_cache"] = cdf
        return cdf


# --------------------------------------------------------------------------------------
# events, rays
# --------------------------------------------------------------------------------------


class Event(enum.Enum):
    """Photon events; ``REACT`` is the fork-specific terminal event of a ``Reactor`` component
    (consumed at reference ``miniplant/simulation_runner.py:66``)."""

    GENERATE = 0
    REFLECT = 1
    TRANSMIT = 2
    ABSORB = 3
    SCATTER = 4
    EMIT = 5
    EXIT = 6
    KILL = 7
    REACT = 8


@dataclass(frozen=True)
class Ray:
    position: tuple
    direction: tuple
    wavelength: float
    source: Optional[str] = None
    travelled: float = 0.0

    def propagate(self, distance: float) -> "Ray":
        p = np.asarray(self.position, dtype=float) + distance * np.asarray(self.direction, dtype=float)
        return replace(self, position=tuple(p.tolist()), travelled=self.travelled + distance)

    def representation(self, from_node: "Node", to_node: "Node") -> "Ray":
        return replace(
            self,
            position=from_node.point_to_node(self.position, to_node),
            direction=from_node.vector_to_node(self.direction, to_node),
        )


# --------------------------------------------------------------------------------------
# materials
# --------------------------------------------------------------------------------------


class Surface:
    """Default pvtrace surface: unpolarised Fresnel reflection / Snell refraction.  Data marker only."""

    kind = "fresnel"


class Component:
    def __init__(self, name="Component"):
        self.name = name


class Scatterer(Component):
    """Base of absorbing components: constant or tabulated (N,2) coefficient in m^-1."""

    def __init__(self, coefficient, x=None, quantum_yield=1.0, phase_function=None, hist=False, name="Scatterer"):
        super().__init__(name=name)
        self._coefficient = coefficient
        self._quantum_yield = quantum_yield
        self._abs_dist = None
        if not np.isscalar(coefficient):
            coefficient = np.asarray(coefficient, dtype=float)
            if coefficient.ndim != 2 or coefficient.shape[1] != 2:
                raise ValueError("tabulated coefficient must be an (N, 2) array")
            self._coefficient = coefficient
            self._abs_dist = Distribution(coefficient[:, 0], coefficient[:, 1], hist=hist)
        self.phase_function = phase_function if phase_function is not None else isotropic

    @property
    def quantum_yield(self):
        return self._quantum_yield

    def coefficient(self, wavelength):
        if self._abs_dist is None:
            return float(self._coefficient)
        return float(self._abs_dist(wavelength))


class Absorber(Scatterer):
    def __init__(self, coefficient, x=None, name="Absorber", hist=False):
        super().__init__(coefficient, x=x, quantum_yield=0.0, phase_function=None, hist=hist, name=name)


class Reactor(Absorber):
    """Fork-specific absorber whose absorption terminates the photon with ``Event.REACT``
    (reference ``miniplant/scene_creator.py:216``)."""

    def __init__(self, coefficient, x=None, name="Reactor", hist=False):
        super().__init__(coefficient, x=x, name=name, hist=hist)


class Luminophore(Scatterer):
    def __init__(self, coefficient, emission=None, x=None, hist=False, quantum_yield=1.0,
                 phase_function=None, name="Luminophore"):
        super().__init__(coefficient, x=x, quantum_yield=quantum_yield, phase_function=phase_function,
                         hist=hist, name=name)
        if emission is None:
            raise ValueError("Luminophore needs an (N, 2) emission spectrum")
        emission = np.asarray(emission, dtype=float)
        self._emission = emission
        self._ems_dist = Distribution(emission[:, 0], emission[:, 1], hist=hist)


class Material:
    """Refractive index + list of absorbing components.  The fork accepts visual keyword
    arguments (``color``, ``transparent``, ``opacity``, ``reflectivity``) that do not affect
    tracing (reference ``miniplant/scene_creator.py:104-110``)."""

    def __init__(self, refractive_index, surface=None, components=None, **visual):
        self.refractive_index = float(refractive_index)
        self.surface = surface if surface is not None else Surface()
        self.components = list(components) if components is not None else []
        self.visual = dict(visual)
        for key, value in visual.items():
            setattr(self, key, value)


# --------------------------------------------------------------------------------------
# geometry (pure descriptions)
# --------------------------------------------------------------------------------------


class Geometry:
    def __init__(self, material=None):
        self.material = material
        self.color = None
        self.transparency = None
        self.opacity = None


class Box(Geometry):
    """Axis-aligned box centred on the local origin, faces at +-size/2."""

    def __init__(self, size, material=None):
        super().__init__(material)
        self.size = tuple(float(s) for s in size)


class Sphere(Geometry):
    def __init__(self, radius, material=None):
        super().__init__(material)
        self.radius = float(radius)


class Cylinder(Geometry):
    """Finite cylinder, axis = local z, centred, flat caps at +-length/2."""

    def __init__(self, length, radius, material=None):
        super().__init__(material)
        self.length = float(length)
        self.radius = float(radius)


# --------------------------------------------------------------------------------------
# lights
# --------------------------------------------------------------------------------------


class Light:
    """Callable-driven photon source (wavelength in nm, position, direction in the light's frame)."""

    def __init__(self, wavelength: Callable = None, position: Callable = None,
                 direction: Callable = None, name: str = "Light"):
        self.wavelength = wavelength if wavelength is not None else (lambda: 555.0)
        self.position = position if position is not None else (lambda: (0.0, 0.0, 0.0))
        self.direction = direction if direction is not None else (lambda: (0.0, 0.0, 1.0))
        self.name = name

    def emit(self, num_rays=None) -> Iterator[Ray]:
        count = 0
        while True:
            count += 1
            if num_rays is not None and count > num_rays:
                break
            yield Ray(
                wavelength=self.wavelength(),
                position=self.position(),
                direction=self.direction(),
                source=self.name,
            )


# --------------------------------------------------------------------------------------
# scene graph
# --------------------------------------------------------------------------------------


class Node(NodeMixin):
    """Scene-graph node: optional geometry, optional light, 4x4 pose relative to its parent."""

    def __init__(self, name=None, parent=None, location=None, geometry=None, light=None):
        self.name = name
        self.geometry = geometry
        self.light = light
        self.pose = np.identity(4)
        if location is not None:
            self.translate(location)
        self.parent = parent

    def __repr__(self):
        return f"Node({self.name!r})"

    # transforms act in the parent frame (pre-multiplication)
    def translate(self, delta):
        self.pose = translation_matrix(delta) @ self.pose

    def rotate(self, angle, axis):
        self.pose = rotation_matrix(angle, axis) @ self.pose

    @property
    def location(self):
        return tuple(self.pose[:3, 3].tolist())

    def world_pose(self) -> np.ndarray:
        """Local -> root transform (product of poses along the path from the root)."""
        m = np.identity(4)
        for node in self.path:
            m = m @ np.asarray(getattr(node, "pose", np.identity(4)), dtype=float)
        return m

    def transformation_to(self, node: "Node") -> np.ndarray:
        """Matrix taking coordinates in this node's frame to ``node``'s frame."""
        if self is node:
            return np.identity(4)
        return np.linalg.inv(node.world_pose()) @ self.world_pose()

    def point_to_node(self, point, node) -> tuple:
        h = np.ones(4)
        h[:3] = np.asarray(point, dtype=float)[:3]
        return tuple((self.transformation_to(node) @ h)[:3].tolist())

    def vector_to_node(self, vector, node) -> tuple:
        m = self.transformation_to(node)[:3, :3]
        return tuple((m @ np.asarray(vector, dtype=float)[:3]).tolist())

    def emit(self, num_rays=None) -> Iterator[Ray]:
        if self.light is None:
            raise ValueError("Node has no light to emit from")
        for ray in self.light.emit(num_rays):
            yield ray.representation(self, self.root)


class Scene:
    def __init__(self, root=None):
        self.root = root

    @property
    def light_nodes(self) -> list:
        found = []
        for node in LevelOrderIter(self.root):
            if isinstance(getattr(node, "light", None), Light):
                found.append(node)
        return found

    @property
    def component_nodes(self) -> list:
        return [n for n in LevelOrderIter(self.root) if getattr(n, "geometry", None) is not None]

    def emit(self, num_rays) -> Iterator[Ray]:
        """Rays in the root frame; rays are dealt round-robin over the light nodes."""
        lights = self.light_nodes
        if not lights or not num_rays:
            return
        root = self.root
        streams = [self._emit_from(node, root) for node in lights]
        for idx in range(int(num_rays)):
            yield next(streams[idx % len(streams)])

    @staticmethod
    def _emit_from(node, root) -> Iterator[Ray]:
        world = _pose_to_root(node)
        for ray in node.light.emit(None):
            h = np.ones(4)
            h[:3] = np.asarray(ray.position, dtype=float)[:3]
            pos = tuple((world @ h)[:3].tolist())
            vec = tuple((world[:3, :3] @ np.asarray(ray.direction, dtype=float)[:3]).tolist())
            yield replace(ray, position=pos, direction=vec)

    def simulate(self, num_rays, workers=None, **kwargs):
        """pvtrace's ``Scene.simulate`` is a CPU process pool (reference call site
        ``miniplant/simulation_runner.py:62``).  The B200 build has no CPU tracing path: the
        product entry point ``miniplant.simulation_runner._common_simulation_runner`` routes every
        photon through the CUDA library, so this method delegates there and returns the terminal
        events only, one ``[(None, Event)]`` history per photon."""
        from ..backend import simulate_final_events

        return simulate_final_events(self, int(num_rays), **kwargs)


def _pose_to_root(node) -> np.ndarray:
    m = np.identity(4)
    for n in node.path:
        m = m @ np.asarray(getattr(n, "pose", np.identity(4)), dtype=float)
    return m


class MeshcatRenderer:
    """Visualisation is out of the hot path; constructing the renderer
    fails loudly instead of silently doing nothing."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError(
            "MeshcatRenderer is not available in the B200 build (meshcat is a visualisation-only "
            "dependency of the reference; see DESIGN.md, out of scope)."
        )
