"""Run a direct- or diffuse-light simulation — mirror of reference ``miniplant/simulation_runner.py``.

Same entry points, arguments, return values and error behaviour; the per-photon Python loop
(``for ray in scene.emit(n): photon_tracer.follow(scene, ray)``, reference ``:38-41``, or
``scene.simulate(num_rays, workers)``, ``:62-63``) is replaced by one launch of the persistent CUDA
photon kernel through the C-ABI of ``include/lscpm.h``.  Keyword-only extras (``seed``, ``device``,
``return_tallies``) default to the reference's behaviour: unseeded, one float returned.
"""
from __future__ import annotations

import logging
from typing import Callable

import numpy as np

from ..backend import capture_paths, simulate_scene
from ..compat import pvtrace_namespace
from .scene_creator import create_direct_scene, create_diffuse_scene

_pv = pvtrace_namespace()
Scene, Event, MeshcatRenderer = _pv.Scene, _pv.Event, _pv.MeshcatRenderer

logger = logging.getLogger("pvtrace").getChild("miniplant")

PV_KEYWORDS = ("add_bottom_PV", "add_side_PV")


def _common_simulation_runner(
    scene: Scene,
    num_photons: int = 100,
    render: bool = False,
    workers: int = 1,
    *,
    seed: int = None,
    device=None,
    return_tallies: bool = False,
    capture: int = 0,
):
    """Trace ``num_photons`` photons through ``scene`` and return the fraction absorbed by the reaction
    mixture, ``count[Event.REACT] / num_photons`` (reference ``:65-66``).

    ``workers`` selected a CPU process pool in the reference; on the GPU every photon of the call is traced
    by one kernel launch, so the value only takes part in the reference's argument check."""
    logger.debug(f"Starting ray-tracing with {num_photons} photons (Render is {render})")

    if render and workers > 1:
        raise RuntimeError("Sorry, cannot use renderer if more than 1 worker is used!")
    if render:
        # raises: the Meshcat visualiser is outside the hot path (out of scope)
        MeshcatRenderer(open_browser=True)

    result = simulate_scene(scene, num_photons, seed=seed, device=device)
    return _report(result, num_photons, lambda: scene, return_tallies, capture, seed, device)


def _report(result, num_photons, scene_factory, return_tallies=False, capture=0, seed=None, device=None):
    """The reference's result handling (``:65-80``) for a tally of terminal classifications."""
    reacted_fraction = result.reacted_fraction
    logger.debug(f"*** SIMULATION ENDED *** (Efficiency was {reacted_fraction:.3f})")

    # The reference asks next_hit(scene, last_ray) which node the final ray is in / heading to and counts the PV names
    # (:45-53).  A final ray points at a PV node exactly when the photon was absorbed inside it (an absorbed photon's
    # last ray sits inside the opaque box; every other terminal ray faces the plate, a capillary or nothing).
    counts = result.nonzero()
    bottomPV_count = sum(c for name, c in counts.items() if name.startswith("absorb/bottomPV/"))
    sidePV_count = sum(c for name, c in counts.items()
                       if name.split("/")[0] == "absorb" and name.split("/")[1] in ("sidePV1", "sidePV2", "sidePV3", "sidePV4"))
    if bottomPV_count > 0:
        logger.info(f"Bottom PV absorbed {bottomPV_count} photons (i.e. {bottomPV_count/num_photons * 100 :.2f} %) ")
    if sidePV_count > 0:
        logger.info(f"The side PV absorbed {sidePV_count} photons (i.e. {sidePV_count / num_photons * 100 :.2f} %) ")
    if return_tallies:
        return reacted_fraction, result
    # same polymorphic return as the reference (:77-80).  A bulk launch keeps no histories; `capture=N` follows N
    # rays of the same light with per-step records for the photon_path slot (reference :55).
    if bottomPV_count > 0 or sidePV_count > 0:
        scene = scene_factory()
        photon_path = capture_paths(scene, capture, seed=seed, device=device) if capture else []
        return reacted_fraction, scene, photon_path, bottomPV_count, sidePV_count
    return reacted_fraction


def _fast_simulation(kind, tilt_angle, solar_spectrum_function, num_photons, include_dye, scene_kwargs, runner_kwargs,
                     scene_factory, solar_elevation=None, solar_azimuth=None):
    """``run_*_simulation`` without building the scene graph: the plate's flattened geometry is cached per
    (tilt, dye, PV options) and the light descriptor is written directly (``lights.py`` restates the light half of
    ``create_direct_scene`` / ``create_diffuse_scene``).  The reference's drivers call these entry points once per time
    point with ~100 photons (``full_simulation.py:77-97``), where rebuilding 35 nodes cost ten times the trace.
    Returns None when the wavelength callable is not one the device can sample (the caller then takes the scene path,
    which emits on the host)."""
    from .. import backend, flatten, lights, tally
    from .full_simulation import plate_geometry

    if runner_kwargs.get("capture"):
        return None
    light = (lights.direct_light(tilt_angle, solar_elevation, solar_azimuth) if kind == "direct"
             else lights.diffuse_light(tilt_angle))
    flatten._wavelength_source(solar_spectrum_function, light)
    if light.wavelength_callable is not None:
        return None
    try:
        arrays = plate_geometry(tilt_angle, include_dye, **scene_kwargs)
    except TypeError:       # unhashable scene option: no cache key
        return None
    num_photons = int(num_photons)
    seed, device = runner_kwargs.get("seed"), runner_kwargs.get("device")
    if not hasattr(arrays, "bin_name_list"):
        arrays.bin_name_list = arrays.bin_names()
    names = arrays.bin_name_list
    if num_photons <= 0:
        result = tally.TallyResult(names, np.zeros(len(names), np.int64), 0, 0)
    else:
        row = backend.trace_lights(arrays, [light], num_photons, seed=seed, device=device)[0]
        result = tally.TallyResult(names, row[:len(names)].copy(), int(row[len(names)]), int(row[len(names) + 1]))
    return _report(result, num_photons, scene_factory, runner_kwargs.get("return_tallies", False), 0, seed, device)


def run_direct_simulation(
    tilt_angle: int = 0,
    solar_elevation: int = 30,
    solar_azimuth: int = 180,
    solar_spectrum_function: Callable = lambda: 555,
    num_photons: int = 100,
    render: bool = False,
    workers: int = 1,
    include_dye: bool = None,
    **kwargs,
):
    """Create a scene for direct irradiation with the provided parameters and run a simulation on it
    (reference ``:83-105``)."""
    runner_kwargs = {k: kwargs.pop(k) for k in ("seed", "device", "return_tallies", "capture") if k in kwargs}

    def scene_factory():
        return create_direct_scene(tilt_angle=tilt_angle, solar_elevation=solar_elevation, solar_azimuth=solar_azimuth,
                                   solar_spectrum_function=solar_spectrum_function, include_dye=include_dye, **kwargs)

    if render and workers > 1:
        raise RuntimeError("Sorry, cannot use renderer if more than 1 worker is used!")
    if not render:
        logger.debug(f"Starting ray-tracing with {num_photons} photons (Render is {render})")
        fast = _fast_simulation("direct", tilt_angle, solar_spectrum_function, num_photons, include_dye, kwargs, runner_kwargs,
                                scene_factory, solar_elevation, solar_azimuth)
        if fast is not None:
            return fast
    scene = create_direct_scene(
        tilt_angle=tilt_angle,
        solar_elevation=solar_elevation,
        solar_azimuth=solar_azimuth,
        solar_spectrum_function=solar_spectrum_function,
        include_dye=include_dye,
        **kwargs,
    )
    return _common_simulation_runner(scene, num_photons, render, workers, **runner_kwargs)


def run_diffuse_simulation(
    tilt_angle: int = 0,
    solar_spectrum_function: Callable = lambda: 555,
    num_photons: int = 100,
    render: bool = False,
    workers: int = 1,
    include_dye: bool = None,
    **runner_kwargs,
):
    """Create a scene for diffuse irradiation with the provided parameters and run a simulation on it
    (reference ``:108-124``)."""
    def scene_factory():
        return create_diffuse_scene(tilt_angle=tilt_angle, solar_spectrum_function=solar_spectrum_function, include_dye=include_dye)

    if render and workers > 1:
        raise RuntimeError("Sorry, cannot use renderer if more than 1 worker is used!")
    if not render:
        logger.debug(f"Starting ray-tracing with {num_photons} photons (Render is {render})")
        fast = _fast_simulation("diffuse", tilt_angle, solar_spectrum_function, num_photons, include_dye, {}, runner_kwargs,
                                scene_factory)
        if fast is not None:
            return fast
    scene = scene_factory()
    return _common_simulation_runner(scene, num_photons, render, workers, **runner_kwargs)


if __name__ == "__main__":   # reference :127-143, without the Meshcat window (render=True needs the renderer)
    res = run_direct_simulation(
        tilt_angle=40,
        solar_elevation=50,
        solar_azimuth=180,
        add_bottom_PV=True,
        add_side_PV=True,
        num_photons=100,
    )

    bottom = res[-2]
    side = res[-1]
    print(f"SIMU ENDED - bottom={bottom} side={side}")
