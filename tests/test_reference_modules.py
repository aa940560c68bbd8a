"""The drop-in boundary, tested with the REFERENCE'S OWN modules.

``compat.install()`` provides the third-party names (pvtrace / anytree / pvlib) the reference imports; then
``/root/reference/miniplant/{utils,scene_creator,simulation_runner}.py`` are imported **unmodified** (the mirror's
``miniplant`` entries in ``sys.modules`` are set aside for the duration, so ``miniplant.*`` imports inside them resolve
to the reference's own files) and their scenes are flattened / traced by this repo's backend.

``/root/reference`` exists only in the authoring container.  On the GPU box the same assertions run on
*reference-shaped* light objects defined below: classes that expose exactly what the reference's expose
(``LightPosition``: ``tilt_angle`` + ``__call__``, reference ``miniplant/utils.py:87-99``; ``IsotropicPhotonGenerator``:
``tilt_angle`` + ``base_position_generator`` + ``__call__``, ``utils.py:125-142``) and nothing the mirror adds.
"""
import importlib
import importlib.util
import os
import sys

import numpy as np
import pytest

REFERENCE = "/root/reference"
HAVE_REFERENCE = os.path.isfile(os.path.join(REFERENCE, "miniplant", "scene_creator.py"))


@pytest.fixture(scope="module")
def reference_modules():
    """The reference's miniplant modules, imported unmodified from where they lie."""
    if not HAVE_REFERENCE:
        pytest.skip("/root/reference is not present on this machine")
    import lscpm_b200

    lscpm_b200.install()
    saved = {k: v for k, v in sys.modules.items() if k == "miniplant" or k.startswith("miniplant.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        mods = {name: importlib.import_module(f"miniplant.{name}") for name in ("utils", "scene_creator", "simulation_runner")}
        for m in mods.values():
            assert m.__file__.startswith(REFERENCE), m.__file__
        yield mods
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k == "miniplant" or k.startswith("miniplant.")]:
            del sys.modules[k]
        sys.modules.update(saved)


# ---- reference-shaped stand-ins (what travels to the GPU box) --------------------------------------------------------


class LightPosition:
    """Only ``tilt_angle`` and ``__call__`` — as in reference utils.py:87-99 (no ``matrix`` method)."""

    def __init__(self, tilt_angle):
        self.tilt_angle = tilt_angle

    def __call__(self, *args, **kwargs):
        t = np.radians(self.tilt_angle)
        u, v = np.random.uniform(-0.235, 0.235, 2)
        return (np.cos(t) * u, v, -np.sin(t) * u)


class VectorInverter:
    def __init__(self, vector):
        self.vector = vector

    def __call__(self, *args, **kwargs):
        return tuple(-c for c in self.vector)


class IsotropicPhotonGenerator:
    def __init__(self, tilt_angle):
        self.tilt_angle = tilt_angle
        self.base_position_generator = LightPosition(tilt_angle)

    def __call__(self, *args, **kwargs):
        from lscpm_b200.miniplant.utils import create_diffuse_photon

        sky = create_diffuse_photon(self.tilt_angle)
        return np.asarray(self.base_position_generator()) + sky, tuple(-c for c in sky)


def reference_shaped_scene(kind, tilt=30.0, elevation=40.0, azimuth=170.0, wavelength=lambda: 555, **kwargs):
    from lscpm_b200.compat import pvtrace_api as pv
    from lscpm_b200.miniplant import scene_creator as mirror
    from lscpm_b200.miniplant.utils import MyLight

    if kind == "direct":
        sun = pv.spherical_to_cart(np.deg2rad(90 - elevation), np.deg2rad(180 - azimuth))
        node = pv.Node(name="Solar Light", parent=None,
                       light=pv.Light(wavelength=wavelength, direction=VectorInverter(sun), position=LightPosition(tilt)))
        node.translate(sun)
    else:
        node = pv.Node(name="Solar Light", parent=None,
                       light=MyLight(wavelength=wavelength, position_and_direction=IsotropicPhotonGenerator(tilt)))
    return mirror._create_scene_common(tilt_angle=tilt, light_source=node, include_dye=True, **kwargs)


def _mirror_scene(kind, tilt=30.0, elevation=40.0, azimuth=170.0, **kwargs):
    from lscpm_b200.miniplant import scene_creator as mirror

    if kind == "direct":
        return mirror.create_direct_scene(tilt_angle=tilt, solar_elevation=elevation, solar_azimuth=azimuth, **kwargs)
    return mirror.create_diffuse_scene(tilt_angle=tilt)


def _same_flattening(a, b):
    assert a.node_name == b.node_name
    for field in ("node_geom", "node_parent", "node_material", "node_size", "node_pose", "mat_refractive_index",
                  "mat_comp_begin", "comp_kind", "comp_coefficient", "comp_abs_table", "comp_ems_table",
                  "comp_quantum_yield", "table_begin", "table_x", "table_y"):
        assert np.array_equal(getattr(a, field), getattr(b, field)), field
    la, lb = a.light, b.light
    assert la.kind == lb.kind and la.tilt_deg == lb.tilt_deg and la.wavelength_nm == lb.wavelength_nm
    assert la.wavelength_callable is None and lb.wavelength_callable is None
    np.testing.assert_allclose(la.mask_pose, lb.mask_pose, rtol=0, atol=1e-15)
    np.testing.assert_allclose(la.direction, lb.direction, rtol=0, atol=1e-15)
    np.testing.assert_allclose(la.back_off, lb.back_off, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(la.mask_half, lb.mask_half)


# ---- CPU: flattening ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("kind,kwargs", [("direct", {}), ("diffuse", {}),
                                         ("direct", dict(add_bottom_PV=True, add_side_PV=True))])
def test_reference_scene_creator_flattens_like_the_mirror(reference_modules, kind, kwargs):
    from lscpm_b200 import flatten

    sc = reference_modules["scene_creator"]
    if kind == "direct":
        scene = sc.create_direct_scene(tilt_angle=30.0, solar_elevation=40.0, solar_azimuth=170.0, **kwargs)
    else:
        scene = sc.create_diffuse_scene(tilt_angle=30.0)
    light_callable = scene.light_nodes[0].light
    position = getattr(light_callable, "position", None) or light_callable.position_direction.base_position_generator
    assert type(position).__module__ == "miniplant.utils" and not hasattr(position, "matrix")
    _same_flattening(flatten.flatten_scene(scene), flatten.flatten_scene(_mirror_scene(kind, **kwargs)))


def test_reference_photon_factory_is_recognised(reference_modules):
    from lscpm_b200 import flatten
    from lscpm_b200.compat.pvtrace_api import Distribution
    from lscpm_b200.workloads import am15_like_photon_spectrum

    spec = am15_like_photon_spectrum()
    factory = reference_modules["utils"].PhotonFactory(Distribution(spec[:, 0], spec[:, 1]))
    scene = reference_modules["scene_creator"].create_direct_scene(tilt_angle=0, solar_elevation=90,
                                                                   solar_spectrum_function=factory)
    light = flatten.flatten_light(scene)
    assert light.wavelength_callable is None and np.array_equal(light.spectrum, spec)


@pytest.mark.parametrize("kind", ["direct", "diffuse"])
def test_reference_shaped_lights_flatten_like_the_mirror(kind):
    from lscpm_b200 import flatten

    scene = reference_shaped_scene(kind)
    assert not hasattr(LightPosition(0), "matrix")
    _same_flattening(flatten.flatten_scene(scene), flatten.flatten_scene(_mirror_scene(kind)))


def test_reference_light_positions_lie_on_the_flattened_mask(reference_modules):
    """The rays the reference's own callables emit are inside the rectangle the flattener hands to the device."""
    from lscpm_b200 import flatten

    scene = reference_modules["scene_creator"].create_direct_scene(tilt_angle=25.0, solar_elevation=35.0, solar_azimuth=200.0)
    light = flatten.flatten_light(scene)
    rays = list(scene.emit(200))
    rot, origin = light.mask_pose[:, :3], light.mask_pose[:, 3]
    for ray in rays:
        anchor = np.asarray(ray.position) + light.back_off * light.direction  # move down-beam onto the face
        local = rot.T @ (anchor - origin)
        assert abs(local[2]) < 1e-9 and abs(local[0]) <= 0.235 + 1e-12 and abs(local[1]) <= 0.235 + 1e-12
        np.testing.assert_allclose(ray.direction, light.direction, atol=1e-12)


def test_unflattenable_light_becomes_unsupported_scene_error():
    """Objects that miss what the flattener reads raise UnsupportedSceneError (→ host emission), not AttributeError."""
    from lscpm_b200 import flatten
    from lscpm_b200.compat import pvtrace_api as pv
    from lscpm_b200.miniplant import scene_creator as mirror
    from lscpm_b200.miniplant.utils import MyLight

    class IsotropicPhotonGenerator:  # right name, nothing inside
        def __call__(self):
            return (0.0, 0.0, 1.0), (0.0, 0.0, -1.0)

    node = pv.Node(name="Solar Light", parent=None,
                   light=MyLight(wavelength=lambda: 555, position_and_direction=IsotropicPhotonGenerator()))
    scene = mirror._create_scene_common(tilt_angle=0, light_source=node, include_dye=True)
    with pytest.raises(flatten.UnsupportedSceneError):
        flatten.flatten_light(scene)


def test_only_provably_constant_wavelength_callables_go_to_the_device():
    import random

    from lscpm_b200 import flatten
    from lscpm_b200.miniplant.scene_creator import create_direct_scene
    from lscpm_b200.workloads import ConstantWavelength

    def green():
        return 555

    nm = 612.5
    calls = []

    def counted():
        calls.append(1)
        return 555

    for fn, expected in ((green, 555.0), (lambda: 700, 700.0), (lambda: nm, 612.5), (ConstantWavelength(585), 585.0)):
        light = flatten.flatten_light(create_direct_scene(solar_spectrum_function=fn))
        assert light.wavelength_callable is None and light.wavelength_nm == expected
    for fn in (lambda: random.choice([450, 450]), counted, lambda: float(np.random.uniform(500, 500))):
        light = flatten.flatten_light(create_direct_scene(solar_spectrum_function=fn))
        assert light.wavelength_callable is fn
    assert not calls, "an unrecognised callable must not be executed by the flattener"


# ---- GPU: the reference's three bands (reference tests/test_simulation_runner.py:12-33) --------------------------------


def green_photons():
    return 555


def red_photons():
    return 700


@pytest.mark.gpu
def test_reference_shaped_scenes_give_the_reference_bands(built_library):
    from lscpm_b200.miniplant.simulation_runner import _common_simulation_runner

    good = _common_simulation_runner(reference_shaped_scene("direct", 40, 50, 180, green_photons), 200_000, seed=5)
    bad = _common_simulation_runner(reference_shaped_scene("direct", 0, 10, 180, red_photons), 200_000, seed=6)
    standard = _common_simulation_runner(reference_shaped_scene("diffuse", 0, wavelength=green_photons), 200_000, seed=7)
    assert 0.22 <= good <= 0.38 and bad <= 0.15 and 0.22 <= standard <= 0.38, (good, bad, standard)
    # identical to the mirror's own scenes, bit for bit (same flattening, same seed)
    from lscpm_b200.miniplant.simulation_runner import run_direct_simulation

    assert good == run_direct_simulation(tilt_angle=40, solar_elevation=50, solar_azimuth=180,
                                         solar_spectrum_function=green_photons, num_photons=200_000, seed=5)


@pytest.mark.gpu
def test_reference_shaped_pv_scene_returns_the_five_tuple(built_library):
    from lscpm_b200.miniplant.simulation_runner import _common_simulation_runner

    scene = reference_shaped_scene("direct", 40, 50, 180, green_photons, add_bottom_PV=True, add_side_PV=True)
    out = _common_simulation_runner(scene, 50_000, seed=8)
    assert isinstance(out, tuple) and len(out) == 5 and out[3] > 0 and out[4] > 0 and 0.2 <= out[0] <= 0.4


def _pvtrace_style_runner(scene, num_photons, workers):
    """The pvtrace object API exactly as the reference's runner consumes it (reference simulation_runner.py:38-48,62-66),
    written against the names ``compat.install()`` registers under ``pvtrace``: per ray ``scene.emit`` ->
    ``photon_tracer.follow`` -> last event + ``next_hit`` of the last ray; or ``scene.simulate`` for ``workers > 1``."""
    import collections

    from pvtrace import Event, photon_tracer
    from pvtrace.algorithm.photon_tracer import next_hit

    pv_hits = collections.Counter()
    if workers == 1:
        finals = []
        for ray in scene.emit(num_photons):
            history = photon_tracer.follow(scene, ray)
            rays, events = zip(*history)
            finals.append(events[-1])
            ahead = next_hit(scene, history[-1][0])
            if ahead:
                pv_hits[ahead[0].name] += 1
    else:
        finals = [photon[-1][1] for photon in scene.simulate(num_rays=num_photons, workers=workers)]
    return collections.Counter(finals)[Event.REACT] / num_photons, pv_hits


@pytest.mark.gpu
def test_pvtrace_object_api_on_reference_shaped_scenes(built_library):
    """What travels to the GPU box of the next test: the seam the reference's runner uses (emit / follow / next_hit /
    simulate / Event.REACT), on scenes whose lights expose only what the reference's classes expose."""
    bulk, _ = _pvtrace_style_runner(reference_shaped_scene("direct", 40, 50, 180, green_photons), 100_000, workers=2)
    assert 0.22 <= bulk <= 0.38
    per_ray, _ = _pvtrace_style_runner(reference_shaped_scene("direct", 40, 50, 180, green_photons), 600, workers=1)
    assert abs(per_ray - bulk) < 4 * np.sqrt(bulk * (1 - bulk) / 600)
    diffuse, _ = _pvtrace_style_runner(reference_shaped_scene("diffuse", 0, wavelength=green_photons), 100_000, workers=2)
    assert 0.22 <= diffuse <= 0.38
    scene = reference_shaped_scene("direct", 40, 50, 180, green_photons, add_bottom_PV=True, add_side_PV=True)
    _, pv_hits = _pvtrace_style_runner(scene, 400, workers=1)
    assert pv_hits["bottomPV"] > 0 and sum(pv_hits[f"sidePV{i}"] for i in range(1, 5)) > 0


@pytest.mark.gpu
def test_reference_runner_unmodified_through_the_backend(built_library, reference_modules):
    """INTEGRATION.md §3, literally: the reference's simulation_runner + scene_creator + utils, unmodified, traced by
    this repo's CUDA library through the pvtrace-shaped shims (reference tests/test_simulation_runner.py:12-33)."""
    runner = reference_modules["simulation_runner"]
    # workers>1 → scene.simulate (reference :62-63) = one bulk launch
    good = runner.run_direct_simulation(tilt_angle=40, solar_elevation=50, solar_azimuth=180,
                                        solar_spectrum_function=green_photons, num_photons=100_000, workers=2)
    bad = runner.run_direct_simulation(tilt_angle=0, solar_elevation=10, solar_azimuth=180,
                                       solar_spectrum_function=red_photons, num_photons=100_000, workers=2)
    standard = runner.run_diffuse_simulation(solar_spectrum_function=green_photons, num_photons=100_000, workers=2)
    assert 0.22 <= good <= 0.38 and bad <= 0.15 and 0.22 <= standard <= 0.38, (good, bad, standard)
    # workers=1 → scene.emit + photon_tracer.follow + next_hit per ray (reference :38-48), as the reference's tests call it
    assert 0.15 <= runner.run_direct_simulation(tilt_angle=40, solar_elevation=50, solar_azimuth=180,
                                                solar_spectrum_function=green_photons, num_photons=200) <= 0.45
    out = runner.run_direct_simulation(tilt_angle=40, solar_elevation=50, solar_azimuth=180, num_photons=300,
                                       add_bottom_PV=True, add_side_PV=True)
    assert isinstance(out, tuple) and len(out) == 5 and out[3] + out[4] > 0
