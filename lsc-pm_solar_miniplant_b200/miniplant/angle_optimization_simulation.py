"""Tilt-angle sweep (direct light only) — mirror of reference ``miniplant/angle_optimization_simulation.py``.

Same entry point and CSV schema (``apparent_elevation, azimuth, simulation_direct, direct_reacted``, reference
``:93-108``); all time points of a tilt angle go into one kernel launch.
"""
from __future__ import annotations

import logging
import time
from pathlib import Path

import numpy as np

from .. import lights as _lights
from .full_simulation import is_rank_zero, trace_time_points, year_batches
from ._csv import write_csv
from .solar_data import solar_data_for_place_and_time
from .utils import PhotonFactory  # noqa: F401

RAYS_PER_SIMULATIONS = 100
INCLUDE_DYE = True

logger = logging.getLogger("pvtrace").getChild("miniplant")


def evaluate_tilt_angle(tilt_angle: int, location, workers: int = None, time_resolution: int = 1800, *, solar_data=None,
                        seed: int = None, device=None, tracer=None, target_file=None):
    """Run a simulation with the given tilt angle/location combination and save results as CSV."""
    logger.info(f"Starting simulation w/ tilt angle {tilt_angle}")
    if solar_data is None:
        solar_data = solar_data_for_place_and_time(location, tilt_angle, time_resolution=time_resolution)

    start_time = time.perf_counter()
    results = solar_data.copy()
    results["simulation_direct"] = 0.0
    results["direct_reacted"] = 0.0
    # points whose spectrum is identically zero (sunrise / sunset) are skipped (reference :65-67)
    valid = np.array([np.count_nonzero(s._y) != 0 for s in solar_data["direct_spectrum"]], dtype=bool)
    for stamp in solar_data.index[~valid]:
        print(f"skipping this point {stamp} due to low spectrum")
    rows = solar_data[valid]
    batches = year_batches(tilt_angle, rows, diffuse=False) if len(rows) else []
    fractions = trace_time_points(tilt_angle, batches, RAYS_PER_SIMULATIONS, INCLUDE_DYE, seed, device, tracer)
    results.loc[valid, "simulation_direct"] = fractions
    results["direct_reacted"] = results["simulation_direct"] * results["direct_irradiance"]
    print(f"Simulation ended in {(time.perf_counter() - start_time) / 60:.1f} minutes!")

    if target_file is None:
        prefix = f"simulation_results/{location.name}/{location.name}"
        if not INCLUDE_DYE:
            prefix += "_no_dye"
        target_file = Path(f"{prefix}_{np.abs(tilt_angle)}deg_results.csv")
    if is_rank_zero():
        target_file = Path(target_file)
        target_file.parent.mkdir(parents=True, exist_ok=True)
        write_csv(results, target_file, columns=("apparent_elevation", "azimuth", "simulation_direct", "direct_reacted"))
    return results


if __name__ == "__main__":   # reference :111-120
    from miniplant.locations import TOWNSVILLE

    site = TOWNSVILLE

    tilt_range = list(range(0, -40, -5))
    for tilt in tilt_range:
        evaluate_tilt_angle(tilt_angle=tilt, location=site, workers=12, time_resolution=1800)
