"""Oracle-free properties of the traced path (GPU): the scene's mirror symmetries.

The plate, its 16 capillaries (parallel to y, equally spaced in x, symmetric about x = 0) and the light masks are mirror
symmetric in x and in y.  A light that is itself symmetric must therefore give tallies that are symmetric — whatever the
restatement of pvtrace's bookkeeping is, so these checks do not depend on ``oracle/``:

* normal incidence (tilt 0, sun at zenith): channel k and channel 15 - k react equally often, the -x and +x (and the -y and
  +y) edge faces leak equally;
* a sun at azimuth 180 + d against one at 180 - d (reference ``scene_creator.py:288-304``: the sun vector's y component
  changes sign): every x-resolved tally agrees between the two runs, and the -y / +y faces (and the capillaries' two end
  caps) swap roles.

Two-sample binomial bands at 4 sigma on 2e7 photons per run (sigma of a per-channel reacted fraction: 3e-5)."""
import numpy as np
import pytest

from helpers import am15_like_photon_spectrum, bins_outside_band, standard_scene

pytestmark = pytest.mark.gpu

N = 20_000_000


def _trace(seed, **scene_kwargs):
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(standard_scene(**scene_kwargs))
    names = arrays.bin_names()
    row = backend.trace_lights(arrays, [arrays.light], N, seed=seed, device=0)[0]
    assert row[:len(names)].sum() == N
    return names, dict(zip(names, (int(c) for c in row[:len(names)])))


def _mirror_x(name: str) -> str:
    """The bin that a reflection x -> -x maps `name` onto: channel k <-> 15 - k, plate face -x <-> +x."""
    parts = name.split("/")
    if len(parts) >= 2 and "_" in parts[1] and parts[1].rsplit("_", 1)[1].isdigit():
        stem, k = parts[1].rsplit("_", 1)
        parts[1] = f"{stem}_{15 - int(k)}"
    elif len(parts) == 3 and parts[2] in ("-x", "+x"):
        parts[2] = "+x" if parts[2] == "-x" else "-x"
    return "/".join(parts)


def _mirror_y(name: str) -> str:
    """y -> -y: plate face -y <-> +y; a capillary's two end caps swap."""
    parts = name.split("/")
    swap = {"-y": "+y", "+y": "-y", "-cap": "+cap", "+cap": "-cap"}
    if len(parts) == 3 and parts[2] in swap:
        parts[2] = swap[parts[2]]
    return "/".join(parts)


def _assert_mirrored(names, a, b, mirror, label):
    mirrored = [mirror(n) for n in names]
    assert sorted(mirrored) == sorted(names), "the mirror map must permute the bins"
    bad = bins_outside_band(names, [a[n] for n in names], [b[m] for m in mirrored], z=4.0)
    assert not bad, f"{label}: {bad}"
    react = [n for n in names if n.startswith("react/")]
    assert len(react) == 16 and sum(a[n] for n in react) > 0.15 * N


def test_normal_incidence_is_mirror_symmetric_in_x_and_y(built_library):
    names, a = _trace(101, tilt=0, elevation=90, azimuth=180, spectrum=am15_like_photon_spectrum())
    _, b = _trace(202, tilt=0, elevation=90, azimuth=180, spectrum=am15_like_photon_spectrum())
    _assert_mirrored(names, a, b, _mirror_x, "x mirror, independent run")
    _assert_mirrored(names, a, b, _mirror_y, "y mirror, independent run")
    # and within ONE run the mirrored halves agree (counts of disjoint bins of one multinomial: same band, conservatively)
    left = sum(a[f"react/Reaction_mixture_{k}/0"] for k in range(8)) if "react/Reaction_mixture_0/0" in a else None
    if left is not None:
        right = sum(a[f"react/Reaction_mixture_{k}/0"] for k in range(8, 16))
        assert abs(left - right) < 4.0 * np.sqrt(left + right)


def test_mirrored_sun_azimuths_give_y_mirrored_tallies(built_library):
    kw = dict(tilt=30, elevation=35, spectrum=am15_like_photon_spectrum())
    names, east = _trace(303, azimuth=180 - 40, **kw)
    _, west = _trace(404, azimuth=180 + 40, **kw)
    _assert_mirrored(names, east, west, _mirror_y, "sun at 140 deg against 220 deg")
    # the beam comes in obliquely in y: the face it heads for must see more light than the face behind it
    lit, shaded = "exit/LSC-PM Reactor 47x47 cm^2/+y", "exit/LSC-PM Reactor 47x47 cm^2/-y"
    if lit in east:
        assert east[lit] != east[shaded]
        assert (east[lit] > east[shaded]) == (west[shaded] > west[lit])


# ---- an analytic known answer through the whole traced path: the absorbing slab -----------------------------------------


def _slab_scene(alpha, n_plate=1.48, thickness=0.008):
    """A bare plate (no dye, no capillaries) of the LSC-PM family: one constant absorber."""
    from lscpm_b200.compat import pvtrace_api as pv
    from lscpm_b200.miniplant.utils import LightPosition, VectorInverter
    from lscpm_b200.workloads import ConstantWavelength

    world = pv.Node(name="World (air)", geometry=pv.Sphere(radius=10.0, material=pv.Material(refractive_index=1.0)))
    sun = pv.spherical_to_cart(0.0, 0.0)
    light = pv.Node(name="Solar Light", light=pv.Light(wavelength=ConstantWavelength(555.0), direction=VectorInverter(sun),
                                                       position=LightPosition(tilt_angle=0.0)), parent=world)
    light.translate(sun)
    plate = pv.Node(name="plate", geometry=pv.Box(size=(0.47, 0.47, thickness), material=pv.Material(
        refractive_index=n_plate, components=[pv.Absorber(coefficient=alpha)])), parent=world)
    plate.translate((0, 0, -0.5 * thickness))
    return pv.Scene(world)


@pytest.mark.parametrize("elevation", [90.0, 50.0, 25.0])
def test_absorbing_slab_matches_the_closed_form(elevation, built_library):
    """Unpolarised Fresnel reflectivity R at both faces (the same by reciprocity), single-pass survival s = exp(-alpha d /
    cos t): reflected at entry R; out of the bottom (1-R)^2 s / (1 - R^2 s^2); out of the top after entering
    (1-R)^2 R s^2 / (1 - R^2 s^2); absorbed: the rest.  No restatement involved - Snell, Fresnel, Beer-Lambert and the
    geometric series.  The light mask is shrunk to the middle of the plate so that no bounce sequence reaches an edge."""
    from lscpm_b200 import backend, flatten, lights

    alpha, n, d = 20.0, 1.48, 0.008
    arrays = flatten.flatten_scene(_slab_scene(alpha, n, d), with_light=False)
    names = arrays.bin_names()
    table = lights.direct_batches(0.0, [elevation], [180.0], wavelength_nm=555.0)
    table.desc["mask_half"] = 0.1
    photons = 20_000_000
    row = backend.trace_lights(arrays, table, photons, seed=2026, device=0)[0]
    got = dict(zip(names, (int(c) for c in row[:len(names)])))
    assert sum(got.values()) == photons

    ti = np.radians(90.0 - elevation)
    tt = np.arcsin(np.sin(ti) / n)
    ci, ct = np.cos(ti), np.cos(tt)
    rs = (ci - n * ct) / (ci + n * ct)
    rp = (ct - n * ci) / (ct + n * ci)
    R = 0.5 * (rs * rs + rp * rp)
    s = np.exp(-alpha * d / ct)
    geo = 1.0 - R * R * s * s
    expect = {"exit/entry_reflect": R, "exit/plate/-z": (1 - R) ** 2 * s / geo, "exit/plate/+z": (1 - R) ** 2 * R * s * s / geo}
    expect["absorb/plate/0"] = 1.0 - sum(expect.values())
    for name, p in expect.items():
        sigma = np.sqrt(p * (1 - p) / photons)
        assert abs(got[name] / photons - p) < 4.5 * sigma + 2e-6, (name, got[name] / photons, p, sigma)
    assert sum(got[k] for k in expect) == photons, {k: v for k, v in got.items() if v and k not in expect}
    # interaction events per photon: the entry, then one per face visit or absorption - expected number of those per
    # photon that entered: sum over passes k of (R s)^k = 1 / (1 - R s)
    events = row[len(names)] / photons
    assert abs(events - (1.0 + (1.0 - R) / (1.0 - R * s))) < 2e-3


# ---- an analytic known answer for the capillary lattice: index-matched black tubes -------------------------------------------


def _black_tube_scene(with_idle_dye, n_plate=1.48, n_tubes=16, pitch=0.03, r_outer=0.0015875):
    """Non-absorbing plate holding index-matched, totally absorbing tubes (no Fresnel event and no refraction at their
    surface).  `with_idle_dye` adds a luminophore that absorbs nothing, which selects the plate-flight kernels."""
    from lscpm_b200.compat import pvtrace_api as pv
    from lscpm_b200.miniplant.utils import LightPosition, VectorInverter
    from lscpm_b200.workloads import ConstantWavelength

    world = pv.Node(name="World (air)", geometry=pv.Sphere(radius=10.0, material=pv.Material(refractive_index=1.0)))
    sun = pv.spherical_to_cart(0.0, 0.0)
    light = pv.Node(name="Solar Light", light=pv.Light(wavelength=ConstantWavelength(555.0), direction=VectorInverter(sun),
                                                       position=LightPosition(tilt_angle=0.0)), parent=world)
    light.translate(sun)
    components = [pv.Absorber(coefficient=0.0)]
    if with_idle_dye:
        grid = np.linspace(300.0, 800.0, 51)
        components.append(pv.Luminophore(coefficient=np.stack([grid, np.zeros_like(grid)], axis=1),
                                         emission=np.stack([grid, np.exp(-((grid - 620.0) / 30.0) ** 2)], axis=1),
                                         quantum_yield=0.95, phase_function=pv.isotropic))
    plate = pv.Node(name="plate", geometry=pv.Box(size=(0.47, 0.47, 0.008), material=pv.Material(
        refractive_index=n_plate, components=components)), parent=world)
    tube_geom = pv.Cylinder(length=0.47, radius=r_outer, material=pv.Material(
        refractive_index=n_plate, components=[pv.Absorber(coefficient=1e7)]))
    x0 = -0.5 * pitch * (n_tubes - 1)
    for k in range(n_tubes):
        tube = pv.Node(name=f"tube_{k}", geometry=tube_geom, parent=plate)
        tube.rotate(np.radians(90), (1, 0, 0))
        tube.translate((x0 + pitch * k, 0, 0))
    plate.translate((0, 0, -0.004))
    return pv.Scene(world)


@pytest.mark.parametrize("with_idle_dye", [False, True])
@pytest.mark.parametrize("elevation,azimuth", [(90.0, 180.0), (40.0, 90.0), (60.0, 270.0)])
def test_black_tubes_shadow_their_geometric_cross_section(elevation, azimuth, with_idle_dye, built_library):
    """Rays that travel in planes x = const (sun at zenith, or due east / west of the plate: no x component) keep their x:
    such a ray is absorbed iff it enters within r_o of a tube axis, on its first way down - the ones reflected at the
    bottom face retrace an x that missed.  With h = 16 * 2 r_o / 0.47 and the Fresnel reflectivity R of the faces:
    tubes (1-R) h, entry reflection R, bottom (1-R)^2 (1-h) / (1-R^2), top (1-R)^2 (1-h) R / (1-R^2); every tube the same
    share.  Both geometry code paths: cell march (no luminophore) and plate flights (a luminophore that absorbs nothing)."""
    from lscpm_b200 import backend, flatten, lights

    n, r_o, width = 1.48, 0.0015875, 0.47
    arrays = flatten.flatten_scene(_black_tube_scene(with_idle_dye), with_light=False)
    names = arrays.bin_names()
    table = lights.direct_batches(0.0, [elevation], [azimuth], wavelength_nm=555.0)
    table.desc["mask_half"] = (0.235, 0.05)          # the whole width in x; the middle in y, away from the tube ends
    photons = 20_000_000
    row = backend.trace_lights(arrays, table, photons, seed=777, device=0)[0]
    got = dict(zip(names, (int(c) for c in row[:len(names)])))
    assert sum(got.values()) == photons
    handle = backend.scene_handle(arrays, 0)
    assert ("flight=1" in handle.kernel_name) == with_idle_dye, handle.kernel_name

    ti = np.radians(90.0 - elevation)
    tt = np.arcsin(np.sin(ti) / n)
    ci, ct = np.cos(ti), np.cos(tt)
    R = 0.5 * (((ci - n * ct) / (ci + n * ct)) ** 2 + ((ct - n * ci) / (ct + n * ci)) ** 2)
    h = 16 * 2 * r_o / width
    expect = {"tubes": (1 - R) * h, "exit/entry_reflect": R, "exit/plate/-z": (1 - R) ** 2 * (1 - h) / (1 - R * R),
              "exit/plate/+z": (1 - R) ** 2 * (1 - h) * R / (1 - R * R)}
    tubes = np.array([got[f"absorb/tube_{k}/0"] for k in range(16)])
    measured = dict(got, tubes=int(tubes.sum()))
    assert abs(sum(expect.values()) - 1.0) < 1e-12
    for name, p in expect.items():
        sigma = np.sqrt(p * (1 - p) / photons)
        assert abs(measured[name] / photons - p) < 4.5 * sigma + 2e-6, (name, measured[name] / photons, p, sigma)
    assert sum(measured[k] for k in expect) == photons, {k: v for k, v in got.items() if v and not k.startswith("absorb/tube") and k not in expect}
    share = tubes.sum() / 16.0
    assert np.all(np.abs(tubes - share) < 4.5 * np.sqrt(share)), tubes


# ---- two geometry code paths against each other: folded plate flights vs explicit bounces ----------------------------------


def _scene_handle_with_env(arrays, **env):
    import os

    from lscpm_b200 import backend

    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return backend.SceneHandle(arrays, 0)       # the choice of kernel and geometry is made at scene creation
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize("case", [dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()),
                                  dict(tilt=35, elevation=30, azimuth=150, wavelength=585.0),
                                  dict(tilt=0, elevation=60, spectrum=am15_like_photon_spectrum(), add_bottom_PV=True, add_side_PV=True)])
def test_folded_flights_agree_with_explicit_bounces(case, built_library):
    """The default path takes a photon in the plastic straight to its next real event through the unfolded lattice of
    capillary images (plate_flight, pool kernel); with LSCPM_PLATE_FLIGHTS=0 the lane kernel marches cell by cell and
    makes every total internal reflection an iteration of its own.  Different geometry code, different draws per photon,
    same physics: every bin within 4.5 sigma at 2e7 photons, the events per photon (which count the folded reflections) to
    1.5e-3."""
    import torch

    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(standard_scene(**case))
    names = arrays.bin_names()
    rows = {}
    for label, env in (("flights", {}), ("bounces", dict(LSCPM_PLATE_FLIGHTS=0, LSCPM_TRACE_KERNEL="lanes"))):
        scene = _scene_handle_with_env(arrays, **env)
        assert ("flight=1" in scene.kernel_name) == (label == "flights"), scene.kernel_name
        plan = backend.Plan(scene, [arrays.light] * 4, list(range(4)))
        tallies = plan.new_tallies()
        plan.trace_into(tallies, N // 4, seed=11 if label == "flights" else 12)
        torch.cuda.synchronize()
        rows[label] = tallies.sum(dim=0).cpu().numpy()
    a, b = rows["flights"], rows["bounces"]
    assert a[:len(names)].sum() == b[:len(names)].sum() == N
    assert not bins_outside_band(names, a[:len(names)], b[:len(names)], z=4.5)
    assert abs(a[len(names)] / b[len(names)] - 1.0) < 1.5e-3, (a[len(names)] / N, b[len(names)] / N)
