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
from ._csv import write_csv
from .solar_data import solar_data_for_place_and_time
from .utils import PhotonFactory  # noqa: F401  (the reference defines a duplicate here, :38-45)

logger = logging.getLogger("pvtrace").getChild("miniplant")


_geometry_cache: dict = {}


def plate_geometry(tilt_angle, include_dye=True, **scene_kwargs):
    """Flattened geometry (no light) of the scene ``create_direct_scene`` builds for this tilt, cached: the drivers
    ask for the same plate thousands of times (reference ``simulation_runner.py:97-105`` rebuilds it per call)."""
    from .. import flatten

    key = (float(tilt_angle), bool(include_dye) if include_dye is not None else True, tuple(sorted(scene_kwargs.items())))
    arrays = _geometry_cache.get(key)
    if arrays is None:
        if len(_geometry_cache) > 256:
            _geometry_cache.clear()
        scene = create_direct_scene(tilt_angle=tilt_angle, include_dye=include_dye, **scene_kwargs)
        arrays = _geometry_cache[key] = flatten.flatten_scene(scene, with_light=False)
        arrays.react_bins = np.array([name.startswith("react/") for name in arrays.bin_names()])
    return arrays


def spectrum_columns(column):
    """(x, flux[n, N]) of a column of ``Distribution`` objects when they all share their abscissae, else None."""
    first = column.iloc[0] if hasattr(column, "iloc") else column[0]
    x = np.asarray(first._x, dtype=float)
    rows = [d._y for d in column]
    if any(len(y) != len(x) for y in rows):
        return None
    if any(d._x is not first._x and not np.array_equal(d._x, x) for d in column):
        return None
    return x, np.asarray(rows, dtype=float)


def agreed_seed(seed):
    """The seed of a launch: the caller's, or fresh entropy — drawn on rank 0 and broadcast under ``torch.distributed``
    so that every rank keys its photons with the same Philox key (bit-identical tallies for any world size)."""
    from .. import backend

    if seed is not None:
        return int(seed)
    seed = backend.fresh_seed()
    try:
        import torch.distributed as dist
    except Exception:   # pragma: no cover - torch is part of the image
        return seed
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [seed]
        dist.broadcast_object_list(box, src=0)
        seed = int(box[0])
    return seed


def is_rank_zero() -> bool:
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank() == 0
    except Exception:   # pragma: no cover
        pass
    import os

    return int(os.environ.get("RANK", "0")) == 0


def trace_time_points(tilt_angle, light_list, num_photons, include_dye=True, seed=None, device=None, tracer=None):
    """Reacted fraction of every light configuration (a ``lights.BatchTable`` or a list of ``LightArrays``): one
    launch for all of them.

    ``tracer(arrays, lights, photons_per_batch, seed) -> int64 [n, n_bins + 2]`` defaults to the CUDA path
    (sharded over the process group when one is initialised)."""
    from .. import backend

    arrays = plate_geometry(tilt_angle, include_dye)
    if not len(light_list):
        return np.zeros(0)
    seed = agreed_seed(seed)
    if tracer is None:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .. import distributed

            rows = distributed.trace_lights_distributed(arrays, light_list, num_photons, seed).cpu().numpy()
        else:
            rows = backend.trace_lights(arrays, light_list, num_photons, seed=seed, device=device)
    else:
        rows = np.asarray(tracer(arrays, light_list, num_photons, seed))
    return rows[:, :len(arrays.react_bins)][:, arrays.react_bins].sum(axis=1) / float(num_photons)


def year_batches(tilt_angle, solar_data, diffuse: bool = True):
    """Batch table of a solar-data frame: one direct batch per row, then (``diffuse``) one sky-diffuse batch per row.
    Built column-wise when the rows' distributions share their abscissae (the solar stage's always do)."""
    direct = spectrum_columns(solar_data["direct_spectrum"])
    sky = spectrum_columns(solar_data["diffuse_spectrum"]) if diffuse else None
    elevation, azimuth = solar_data["apparent_elevation"].to_numpy(), solar_data["azimuth"].to_numpy()
    if direct is not None and (not diffuse or sky is not None):
        tables = [_lights.direct_batches(tilt_angle, elevation, azimuth, flux=direct[1], x=direct[0])]
        if diffuse:
            tables.append(_lights.diffuse_batches(tilt_angle, flux=sky[1], x=sky[0]))
        return _lights.BatchTable.concatenate(tables)
    batches = [_lights.direct_light(tilt_angle, e, a, s) for e, a, s in zip(elevation, azimuth, solar_data["direct_spectrum"])]
    if diffuse:
        batches += [_lights.diffuse_light(tilt_angle, s) for s in solar_data["diffuse_spectrum"]]
    return batches


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
    n = len(solar_data)
    fractions = trace_time_points(tilt_angle, year_batches(tilt_angle, solar_data), num_photons_per_simulation, include_dye,
                                  seed, device, tracer) if n else np.zeros(0)
    results = solar_data.copy()
    results["simulation_direct"] = fractions[:n]
    results["direct_reacted"] = results["simulation_direct"] * results["direct_irradiance"] * REACTOR_AREA_IN_M2
    results["simulation_diffuse"] = fractions[n:]
    results["diffuse_reacted"] = results["simulation_diffuse"] * results["diffuse_irradiance"] * REACTOR_AREA_IN_M2
    print(f"Simulation ended in {(time.time() - start_time) / 60:.1f} minutes!")

    if not target_file:
        target_file = Path(f"full_simulation_results/{location.name}/{location.name}_{tilt_angle}deg_results.csv")
    if is_rank_zero():   # every rank holds the reduced results; one of them writes the file
        target_file = Path(target_file)
        target_file.parent.mkdir(parents=True, exist_ok=True)
        write_csv(results, target_file, columns=("apparent_elevation", "azimuth", "simulation_direct", "direct_reacted",
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
