"""Year-long simulation with direct and diffuse irradiation — mirror of reference ``miniplant/full_simulation.py``.

Same entry point, arguments and CSV schema (``apparent_elevation, azimuth, simulation_direct, direct_reacted,
simulation_diffuse, diffuse_reacted``, reference ``:115-125``).  The reference applies ``run_direct_simulation`` and
``run_diffuse_simulation`` to the DataFrame row by row (``:64-106``: ~16 000 scene rebuilds and tiny simulations per
year); here every time point becomes one batch of ONE kernel launch (direct and diffuse batches together), sharded
over the ranks when run under ``torchrun``.
"""
from __future__ import annotations

import logging
import time
from pathlib import Path

import numpy as np

from .. import lights as _lights
from .scene_creator import REACTOR_AREA_IN_M2, create_direct_scene
from .solar_data import solar_data_for_place_and_time
from .utils import PhotonFactory  # noqa: F401  (the reference defines a duplicate here, :38-45)

logger = logging.getLogger("pvtrace").getChild("miniplant")


def trace_time_points(tilt_angle, light_list, num_photons, include_dye=True, seed=None, device=None, tracer=None):
    """Reacted fraction of every light configuration: one launch for all of them.

    ``tracer(arrays, lights, photons_per_batch, seed) -> int64 [n, n_bins + 2]`` defaults to the CUDA path
    (sharded over the process group when one is initialised)."""
    from .. import backend, flatten

    arrays = flatten.flatten_scene(create_direct_scene(tilt_angle=tilt_angle, include_dye=include_dye), with_light=False)
    if not light_list:
        return np.zeros(0)
    seed = backend.fresh_seed() if seed is None else int(seed)
    if tracer is None:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .. import distributed

            rows = distributed.trace_lights_distributed(arrays, light_list, num_photons, seed).cpu().numpy()
        else:
            rows = backend.trace_lights(arrays, light_list, num_photons, seed=seed, device=device)
    else:
        rows = np.asarray(tracer(arrays, light_list, num_photons, seed))
    names = arrays.bin_names()
    react = np.array([name.startswith("react/") for name in names])
    return rows[:, :len(names)][:, react].sum(axis=1) / float(num_photons)


def yearlong_simulation(
    tilt_angle: int,
    location,
    workers: int = None,
    time_resolution: int = 1800,
    num_photons_per_simulation: int = 120,
    include_dye: bool = True,
    time_range=None,
    target_file=None,
    *,
    solar_data=None,
    seed: int = None,
    device=None,
    tracer=None,
):
    logger.info(f"Starting simulation w/ tilt angle {tilt_angle}")
    if solar_data is None:
        solar_data = solar_data_for_place_and_time(location, tilt_angle, time_resolution)
    if time_range:
        solar_data = solar_data.loc[time_range[0]: time_range[1]]

    start_time = time.time()
    batches = []
    for _, row in solar_data.iterrows():
        batches.append(_lights.direct_light(tilt_angle, row["apparent_elevation"], row["azimuth"], row["direct_spectrum"]))
    for _, row in solar_data.iterrows():
        batches.append(_lights.diffuse_light(tilt_angle, row["diffuse_spectrum"]))
    fractions = trace_time_points(tilt_angle, batches, num_photons_per_simulation, include_dye, seed, device, tracer)
    n = len(solar_data)
    results = solar_data.copy()
    results["simulation_direct"] = fractions[:n]
    results["direct_reacted"] = results["simulation_direct"] * results["direct_irradiance"] * REACTOR_AREA_IN_M2
    results["simulation_diffuse"] = fractions[n:]
    results["diffuse_reacted"] = results["simulation_diffuse"] * results["diffuse_irradiance"] * REACTOR_AREA_IN_M2
    print(f"Simulation ended in {(time.time() - start_time) / 60:.1f} minutes!")

    if not target_file:
        target_file = Path(f"full_simulation_results/{location.name}/{location.name}_{tilt_angle}deg_results.csv")
    target_file = Path(target_file)
    target_file.parent.mkdir(parents=True, exist_ok=True)
    results.to_csv(target_file, columns=("apparent_elevation", "azimuth", "simulation_direct", "direct_reacted",
                                         "simulation_diffuse", "diffuse_reacted"))
    return results


if __name__ == "__main__":   # reference :128-148
    logging.getLogger("pvtrace").setLevel(logging.WARNING)  # use logging.DEBUG for more printouts

    from miniplant.locations import EINDHOVEN

    sim_to_run = [(EINDHOVEN, 40)]

    for sim_params in sim_to_run:
        print(f"Now simulating {sim_params[0].name} at {sim_params[1]} deg tilt angle")
        yearlong_simulation(
            tilt_angle=sim_params[1],
            location=sim_params[0],
            workers=12,
            time_resolution=60 * 30,
            include_dye=True,
        )
