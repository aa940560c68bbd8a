"""CPU ORACLE (test infrastructure, NOT product code) — literal float64 restatement of the
pvtrace photon-tracing algorithm that the reference's hot path drives.

PARITY PIN STATUS: per-ray values are **parity unpinned** against real pvtrace output (no per-ray vectors exist in
the reference and pvtrace cannot be run here); tallies are pinned statistically on the full-year results the
reference committed (``tests/test_solar_stage.py::test_oracle_against_the_reference_yearly_results``, within 1 %,
see DESIGN.md section 5).  The arithmetic of the path
lives in the third-party package ``pvtrace`` (fork ``cambiegroup/pvtrace`` at commit
``22c90761e2ee2a9b9193051f40ee553bd4f527f8``, reference ``requirements.txt:43``), which is neither
vendored under ``/root/reference`` nor installable in the build image (no network).  This file
restates pvtrace 2.1's published algorithm — generic scene graph, brute-force intersection of every
geometry node, container/adjacent bookkeeping, Beer-Lambert free paths, Fresnel surfaces,
luminophore re-emission — and is anchored on what the reference itself pins:

* the reference's call sites (``miniplant/simulation_runner.py:38-66``: ``scene.emit``,
  ``photon_tracer.follow``, ``next_hit``, ``Counter(finals)[Event.REACT] / num_photons``);
* the reference's own statistical test bands (``tests/test_simulation_runner.py:12-33``), checked in
  ``tests/test_oracle_golden.py``;
* analytically derived known-answer vectors, checked in
  ``tests/test_oracle_known_answers.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this
module.  It works directly on the scene-graph objects (nodes with 4x4 poses), NOT on the flattened
arrays the CUDA library consumes, so it also cross-checks the flattener.

Algorithm restated (pvtrace 2.1, ``pvtrace/algorithm/photon_tracer.py`` and friends):

``next_hit``     all intersections of the ray with every geometry node, each computed in the node's
                 local frame, expressed in the root frame and sorted by distance (stable sort, i.e.
                 exact ties keep scene-graph traversal order); intersections at distance ~0 are
                 dropped; ``container`` = ``find_container(intersections)``: the nearest of the nodes hit exactly
                 once (the ray is inside them); ``hit`` = first intersection's node; ``adjacent`` =
                 ``find_container(intersections[1:])`` if ``container is hit`` (what encloses the point just beyond
                 the surface) else ``hit``.  Both details are pinned on reference output, the 41 committed dye tilt
                 sweeps (3.2e7 reference photons): with ``adjacent = intersections[1].hit`` (a
                 photon leaving a tube wall then sees a sibling capillary's index when one lies straight ahead) the
                 scenes react 0.34 % too much (z = -9.7); with the deepest instead of the nearest node as container
                 0.13 % too much (z = -3.6); as restated here 0.9999 of the reference (z = +0.2).
``follow``       loop (<= ``maxsteps``): ``next_hit``; none -> EXIT; hit is the root -> EXIT; sample a free
                 path ``-ln(1-U)/alpha(lambda)`` in the container's material; if shorter than the distance to
                 the surface: absorb there, pick the component ~ alpha_i, radiative (``U < QY``) -> re-emit
                 (isotropic, wavelength by inverse CDF restricted to ``[cdf(lambda_kT), 1]``) else terminal
                 ABSORB / REACT; otherwise move to the surface, reflect with probability ``R_fresnel`` else
                 refract (Snell).

Coincident faces (documented in DESIGN.md): the distance~0 filter and "distances tie" are evaluated with a
tolerance (``ZERO_TOL``) that stands for exact arithmetic.  Real pvtrace applies ``10*DBL_EPSILON`` to
float64 round-off, so the ORDER of tied intersections (the capillary end caps coincide with the plate's
+-y faces, the side PV boxes with its edge faces) is decided by noise: the photon loop draws it uniformly
at random, which is also what the reference's committed results favour (plate-first order: z = +2.1,
random order: z = +0.2); probes without a generator keep scene-graph pre-order.
"""
from __future__ import annotations

import math
from collections import Counter
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

ZERO_TOL = 1e-12  # metres; "distance is zero / distances are equal" under exact arithmetic
KB_EV = 8.617333262145e-5  # Boltzmann constant, eV/K
HC_EV_NM = 1240.8  # pvtrace's rounded h*c in eV*nm
MAXSTEPS = 1000

# event codes shared with the tally naming in lscpm_b200.tally
GENERATE, REFLECT, TRANSMIT, ABSORB, EMIT, EXIT, KILL, REACT = range(8)
EVENT_NAMES = ("GENERATE", "REFLECT", "TRANSMIT", "ABSORB", "EMIT", "EXIT", "KILL", "REACT")

BOX_FACES = ("-x", "+x", "-y", "+y", "-z", "+z")
CYL_FACES = ("side", "-cap", "+cap")


# --------------------------------------------------------------------------------------
# geometry in the node-local frame
# --------------------------------------------------------------------------------------


def _box_hits(half, o, d):
    """Slab intersection of a ray with the centred box |x_i| <= half_i.  Returns [(t, face)] for
    the entry and exit points with t >= 0 (pvtrace keeps both end points of the clipped line)."""
    tmin, tmax = -math.inf, math.inf
    fmin = fmax = -1
    for ax in range(3):
        if d[ax] == 0.0:
            if abs(o[ax]) > half[ax]:
                return []
            continue
        t_lo = (-half[ax] - o[ax]) / d[ax]
        t_hi = (half[ax] - o[ax]) / d[ax]
        f_lo, f_hi = 2 * ax, 2 * ax + 1
        if t_lo > t_hi:
            t_lo, t_hi, f_lo, f_hi = t_hi, t_lo, f_hi, f_lo
        if t_lo > tmin:
            tmin, fmin = t_lo, f_lo
        if t_hi < tmax:
            tmax, fmax = t_hi, f_hi
    if tmin > tmax:
        return []
    hits = []
    if tmin >= 0.0:
        hits.append((tmin, fmin))
    if tmax >= 0.0:
        hits.append((tmax, fmax))
    return hits


def _quadratic(a, b, c):
    if a == 0.0:
        return []
    disc = b * b - 4.0 * a * c
    if disc < 0.0:
        return []
    root = math.sqrt(disc)
    # numerically stable pair
    q = -0.5 * (b + math.copysign(root, b)) if b != 0.0 else 0.5 * root
    if q == 0.0:
        return [0.0, 0.0]
    return sorted((q / a, c / q))


def _cylinder_hits(length, radius, o, d):
    """Finite z-axis cylinder: side roots with |z| <= L/2 plus cap discs; t >= 0."""
    hits = []
    half = 0.5 * length
    a = d[0] * d[0] + d[1] * d[1]
    b = 2.0 * (d[0] * o[0] + d[1] * o[1])
    c = o[0] * o[0] + o[1] * o[1] - radius * radius
    for t in _quadratic(a, b, c):
        if t >= 0.0 and abs(o[2] + t * d[2]) <= half:
            hits.append((t, 0))
    if d[2] != 0.0:
        for face, zc in ((1, -half), (2, half)):
            t = (zc - o[2]) / d[2]
            if t >= 0.0:
                x, y = o[0] + t * d[0], o[1] + t * d[1]
                if x * x + y * y < radius * radius:
                    hits.append((t, face))
    return hits


def _sphere_hits(radius, o, d):
    a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    b = 2.0 * (d[0] * o[0] + d[1] * o[1] + d[2] * o[2])
    c = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] - radius * radius
    return [(t, 0) for t in _quadratic(a, b, c) if t >= 0.0]


def _local_normal(kind, params, face, p_local):
    if kind == "box":
        n = [0.0, 0.0, 0.0]
        n[face // 2] = 1.0 if face % 2 else -1.0
        return np.array(n)
    if kind == "cylinder":
        if face == 0:
            r = math.hypot(p_local[0], p_local[1])
            return np.array([p_local[0] / r, p_local[1] / r, 0.0])
        return np.array([0.0, 0.0, -1.0 if face == 1 else 1.0])
    r = math.sqrt(float(np.dot(p_local, p_local)))
    return np.asarray(p_local, dtype=float) / r


# --------------------------------------------------------------------------------------
# optics (pvtrace/material/utils + surface delegate)
# --------------------------------------------------------------------------------------


def fresnel_reflectivity(angle, n1, n2):
    """Unpolarised Fresnel reflectivity at incidence ``angle`` (radians) going n1 -> n2; 1 above the
    critical angle."""
    if n2 < n1 and angle > math.asin(n2 / n1):
        return 1.0
    c = math.cos(angle)
    s = math.sin(angle)
    k = math.sqrt(max(0.0, 1.0 - (n1 / n2 * s) ** 2))
    rs = ((n1 * c - n2 * k) / (n1 * c + n2 * k)) ** 2
    rp = ((n1 * k - n2 * c) / (n1 * k + n2 * c)) ** 2
    return 0.5 * (rs + rp)


def specular_reflection(direction, normal):
    d = np.asarray(direction, dtype=float)
    n = np.asarray(normal, dtype=float)
    return d - 2.0 * float(np.dot(d, n)) * n


def fresnel_refraction(direction, normal, n1, n2):
    d = np.asarray(direction, dtype=float)
    nrm = np.asarray(normal, dtype=float)
    n = n1 / n2
    dot = float(np.dot(d, nrm))
    c = math.sqrt(max(0.0, 1.0 - n * n * (1.0 - dot * dot)))
    sign = -1.0 if dot < 0.0 else 1.0
    return n * d + sign * (c - sign * n * dot) * nrm


def facing_normal(direction, normal):
    """Normal flipped so that it points along the ray (pvtrace's 'be flexible with the normal')."""
    return -normal if float(np.dot(normal, direction)) < 0.0 else normal


def kT_shifted_wavelength(nm, temperature=300.0):
    ev = HC_EV_NM / nm + 1.5 * KB_EV * temperature
    return HC_EV_NM / ev


# --------------------------------------------------------------------------------------
# scene wrapper
# --------------------------------------------------------------------------------------


@dataclass
class _ONode:
    index: int
    name: str
    kind: str  # box | cylinder | sphere
    params: tuple
    to_world: np.ndarray
    to_local: np.ndarray
    refractive_index: float
    components: list
    source: object = None


@dataclass
class Step:
    position: np.ndarray
    direction: np.ndarray
    wavelength: float
    event: int
    node: Optional[int] = None  # hit node (surface events) or container (absorption events)
    face: Optional[int] = None
    component: Optional[int] = None
    container: Optional[int] = None
    adjacent: Optional[int] = None
    distance: Optional[float] = None
    reflectivity: Optional[float] = None


def _component_kind(comp) -> str:
    names = [c.__name__ for c in type(comp).__mro__]
    if "Reactor" in names:
        return "reactor"
    if "Luminophore" in names:
        return "luminophore"
    return "absorber"


class OracleScene:
    """Pre-processed view of a scene graph (shim or real pvtrace objects, duck-typed)."""

    def __init__(self, scene):
        self.scene = scene
        self.nodes: list[_ONode] = []
        self._walk(scene.root, np.identity(4))
        if not self.nodes or self.nodes[0].source is not scene.root:
            raise ValueError("scene root must carry the world geometry")

    def _walk(self, node, parent_world):
        pose = np.asarray(getattr(node, "pose", np.identity(4)), dtype=float)
        world = parent_world @ pose
        geometry = getattr(node, "geometry", None)
        if geometry is not None:
            cls = type(geometry).__name__
            if cls == "Box":
                kind, params = "box", tuple(0.5 * float(s) for s in geometry.size)
            elif cls == "Cylinder":
                kind, params = "cylinder", (float(geometry.length), float(geometry.radius))
            elif cls == "Sphere":
                kind, params = "sphere", (float(geometry.radius),)
            else:
                raise ValueError(f"unsupported geometry {cls}")
            material = geometry.material
            self.nodes.append(
                _ONode(
                    index=len(self.nodes), name=str(node.name), kind=kind, params=params,
                    to_world=world, to_local=np.linalg.inv(world),
                    refractive_index=float(material.refractive_index),
                    components=list(material.components), source=node,
                )
            )
        for child in node.children:
            self._walk(child, world)

    # -- intersections -------------------------------------------------------------
    def intersections(self, position, direction, rng=None):
        """[(distance, node_index, face, point_world)] sorted by distance.  Exact ties (coincident faces): pvtrace sorts
        distances that each node computed in its own frame, so float64 round-off decides; restated as a uniformly
        random order when ``rng`` is given (the photon loop), as scene-graph traversal order otherwise (probes)."""
        position = np.asarray(position, dtype=float)
        direction = np.asarray(direction, dtype=float)
        found = []
        for node in self.nodes:
            o = node.to_local[:3, :3] @ position + node.to_local[:3, 3]
            d = node.to_local[:3, :3] @ direction
            if node.kind == "box":
                hits = _box_hits(node.params, o, d)
            elif node.kind == "cylinder":
                hits = _cylinder_hits(node.params[0], node.params[1], o, d)
            else:
                hits = _sphere_hits(node.params[0], o, d)
            for t, face in hits:
                # poses are rigid, so the local parameter is the world distance for a unit direction
                found.append((t, node.index, face))
        # stable order with tolerance: insertion sort on (distance, traversal order)
        ordered = []
        for item in found:
            pos = len(ordered)
            while pos > 0:
                prev = ordered[pos - 1]
                if prev[0] > item[0] + ZERO_TOL or (abs(prev[0] - item[0]) <= ZERO_TOL and prev[1] > item[1]):
                    pos -= 1
                else:
                    break
            ordered.insert(pos, item)
        if rng is not None:
            start = 0
            for end in range(1, len(ordered) + 1):
                if end == len(ordered) or ordered[end][0] - ordered[start][0] > ZERO_TOL:
                    if end - start > 1:
                        ordered[start:end] = [ordered[start + int(k)] for k in rng.permutation(end - start)]
                    start = end
        return [(t, idx, face, position + t * direction) for t, idx, face in ordered]

    def next_hit(self, position, direction, rng=None):
        """Literal ``next_hit``: returns None or (hit, container, adjacent, point, distance, face)."""
        inter = [x for x in self.intersections(position, direction, rng) if not abs(x[0]) <= ZERO_TOL]
        if not inter:
            return None
        first = inter[0]
        if len(inter) == 1:
            return first[1], first[1], None, first[3], first[0], first[2]
        container = self.find_container(inter)
        hit = first[1]
        adjacent = self.find_container(inter[1:]) if container == hit else hit
        return hit, container, adjacent, first[3], first[0], first[2]

    def find_container(self, inter):
        """pvtrace ``find_container``: the nearest of the nodes that appear exactly once in the list (the ray is inside
        a convex node iff it leaves it once)."""
        if len(inter) == 1:
            return inter[0][1]
        counts = Counter(x[1] for x in inter)
        for x in inter:
            if counts[x[1]] == 1:
                return x[1]
        raise RuntimeError("no container: the ray is outside every node (pvtrace raises here as well)")

    # -- materials -----------------------------------------------------------------
    def coefficients(self, node_index, wavelength):
        return [float(c.coefficient(wavelength)) for c in self.nodes[node_index].components]

    def surface_normal_world(self, node_index, face, point_world):
        node = self.nodes[node_index]
        p_local = node.to_local[:3, :3] @ point_world + node.to_local[:3, 3]
        return node.to_world[:3, :3] @ _local_normal(node.kind, node.params, face, p_local)

    # -- the photon loop -----------------------------------------------------------
    def follow(self, position, direction, wavelength, rng: np.random.Generator, maxsteps=MAXSTEPS,
               emit_method="kT"):
        position = np.asarray(position, dtype=float)
        direction = np.asarray(direction, dtype=float)
        history = [Step(position, direction, wavelength, GENERATE)]
        count = 0
        while True:
            count += 1
            if count > maxsteps:
                history.append(Step(position, direction, wavelength, KILL))
                break
            info = self.next_hit(position, direction, rng)
            if info is None:
                history.append(Step(position, direction, wavelength, EXIT))
                break
            hit, container, adjacent, point, full_distance, face = info
            if hit == 0:
                history.append(Step(point, direction, wavelength, EXIT, node=0, distance=full_distance))
                break
            coefs = self.coefficients(container, wavelength)
            alpha = float(np.sum(coefs)) if coefs else 0.0
            if np.isclose(alpha, 0.0):
                depth = math.inf
            elif not np.isfinite(alpha):
                depth = 0.0
            else:
                depth = -math.log(1.0 - rng.random()) / alpha
            if depth < full_distance:
                position = position + depth * direction
                cdf = np.cumsum(coefs)
                edges = np.hstack([0.0, cdf / cdf.max()])
                index = int(math.floor(np.interp(rng.random(), edges, np.arange(len(coefs) + 1))))
                index = min(index, len(coefs) - 1)
                comp = self.nodes[container].components[index]
                kind = _component_kind(comp)
                radiative = kind == "luminophore" and rng.random() < float(comp._quantum_yield)
                if radiative:
                    phi = rng.uniform(0.0, 2.0 * math.pi)
                    mu = rng.uniform(-1.0, 1.0)
                    theta = math.acos(mu)
                    direction = np.array([math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi),
                                          math.cos(theta)])
                    dist = comp._ems_dist
                    if emit_method == "kT":
                        p1 = float(dist.lookup(kT_shifted_wavelength(wavelength)))
                    elif emit_method == "redshift":
                        p1 = float(dist.lookup(wavelength))
                    else:
                        p1 = 0.0
                    wavelength = float(dist.sample(rng.uniform(p1, 1.0)))
                    history.append(Step(position, direction, wavelength, EMIT, node=container, component=index,
                                        container=container, distance=depth))
                    continue
                event = REACT if kind == "reactor" else ABSORB
                history.append(Step(position, direction, wavelength, event, node=container, component=index,
                                    container=container, distance=depth))
                break
            position = point
            n1 = self.nodes[container].refractive_index
            n2 = self.nodes[adjacent].refractive_index
            normal = facing_normal(direction, self.surface_normal_world(hit, face, point))
            cosang = min(1.0, max(-1.0, float(np.dot(normal, direction))))
            reflectivity = fresnel_reflectivity(math.acos(cosang), n1, n2)
            if rng.random() < reflectivity:
                direction = specular_reflection(direction, normal)
                event = REFLECT
            else:
                direction = fresnel_refraction(direction, normal, n1, n2)
                event = TRANSMIT
            history.append(Step(position, direction, wavelength, event, node=hit, face=face, container=container,
                                adjacent=adjacent, distance=full_distance, reflectivity=reflectivity))
        return history

    # -- deterministic single-interaction probe (parity check #1) --------------------
    def probe(self, position, direction):
        """Deterministic part of one step: next hit, interface, normal, reflectivity and both outgoing
        directions.  Returns None when the ray leaves the scene."""
        info = self.next_hit(position, direction)
        if info is None:
            return None
        hit, container, adjacent, point, distance, face = info
        out = dict(hit=hit, container=container, adjacent=adjacent, point=point, distance=distance, face=face)
        if adjacent is not None and hit != 0:
            direction = np.asarray(direction, dtype=float)
            n1 = self.nodes[container].refractive_index
            n2 = self.nodes[adjacent].refractive_index
            normal = facing_normal(direction, self.surface_normal_world(hit, face, point))
            cosang = min(1.0, max(-1.0, float(np.dot(normal, direction))))
            r = fresnel_reflectivity(math.acos(cosang), n1, n2)
            out.update(n1=n1, n2=n2, normal=normal, reflectivity=r,
                       reflected=specular_reflection(direction, normal),
                       refracted=None if r >= 1.0 else fresnel_refraction(direction, normal, n1, n2))
        return out

    # -- tally naming (shared contract with the CUDA library's bins) -----------------
    def classify(self, history) -> str:
        """Bin name of a finished photon; see lscpm_b200/tally.py for the naming contract."""
        last = history[-1]
        if last.event == KILL:
            return "kill"
        if last.event in (ABSORB, REACT):
            node = self.nodes[last.node]
            prefix = "react" if last.event == REACT else "absorb"
            return f"{prefix}/{node.name}/{last.component}"
        surface = [s for s in history if s.event in (REFLECT, TRANSMIT)]
        if not surface:
            return "exit/miss"
        final = surface[-1]
        if len(surface) == 1 and final.event == REFLECT and final.container == 0 \
                and not any(s.event == EMIT for s in history):
            return "exit/entry_reflect"
        node = self.nodes[final.node]
        faces = BOX_FACES if node.kind == "box" else CYL_FACES if node.kind == "cylinder" else ("surface",)
        return f"exit/{node.name}/{faces[final.face]}"


# --------------------------------------------------------------------------------------
# runner equivalent to miniplant/simulation_runner.py:_common_simulation_runner
# --------------------------------------------------------------------------------------


@dataclass
class OracleResult:
    num_photons: int
    tallies: Counter = field(default_factory=Counter)
    finals: Counter = field(default_factory=Counter)
    events: int = 0  # follow-loop iterations that performed an interaction (surface or absorption)
    emissions: int = 0

    @property
    def reacted_fraction(self) -> float:
        return self.finals[REACT] / self.num_photons if self.num_photons else 0.0


def _seed_global_numpy(seed):
    state = np.random.get_state()
    np.random.seed(seed)
    return state


def run(scene, num_photons: int, seed: int = 0, keep_histories: bool = False) -> OracleResult:
    """Emit ``num_photons`` rays from the scene's light (through the very callables the reference
    binds, driven by the global numpy RNG like the reference) and follow each to termination."""
    oscene = OracleScene(scene)
    rng = np.random.default_rng(seed)
    saved = _seed_global_numpy(seed % (2 ** 32))
    result = OracleResult(num_photons=num_photons)
    histories = []
    try:
        for ray in scene.emit(num_photons):
            history = oscene.follow(ray.position, ray.direction, float(ray.wavelength), rng)
            result.finals[history[-1].event] += 1
            result.tallies[oscene.classify(history)] += 1
            result.events += sum(1 for s in history if s.event in (REFLECT, TRANSMIT, ABSORB, REACT, EMIT))
            result.emissions += sum(1 for s in history if s.event == EMIT)
            if keep_histories:
                histories.append(history)
    finally:
        np.random.set_state(saved)
    if keep_histories:
        result.histories = histories
    return result
