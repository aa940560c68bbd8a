"""Solar irradiance and spectrum incident on the reactor — mirror of reference ``miniplant/solar_data.py``.

Same entry point and output columns (``solar_data_for_place_and_time(site, tilt_angle, time_resolution)`` returns a
DataFrame with ``apparent_elevation``, ``azimuth``, ``aoi``, ``direct_spectrum`` / ``diffuse_spectrum`` photon-flux
``Distribution`` objects on the 27 wavelengths 360-690 nm and ``direct_irradiance`` / ``diffuse_irradiance``), computed
for all time points at once instead of row by row.

pvlib is not installable in the build image, so its three calls are restated in ``compat/pvlib_lite.py`` and checked
against what the reference itself committed:

* **solar position** — ``pvlib_lite.solar_position`` (NOAA/Meeus series + SPA refraction), within 0.015 deg of the
  ephemerides in the reference's result CSVs;
* **air mass** — Kasten & Young (1989), pvlib's default;
* **spectra** — ``pvlib_lite.spectrum.spectrl2``: Bird & Riordan's SPCTRAL2 on the 27 wavelengths the reference keeps.
  The direct and sky-diffuse irradiances recoverable from the reference's result CSVs are reproduced to 2e-5 relative
  when the model is fed the CSVs' own solar positions, to ~1e-3 through the solar position above.

When a real pvlib is importable its own ``spectrl2`` is called, exactly as the reference does.
"""
from __future__ import annotations

import datetime
import logging

import numpy as np
import pandas as pd

from ..compat import pvlib_lite, pvtrace_namespace
from ..workloads import SPECTRL2_SLICE_NM
from ..compat.pvtrace_api import LazyDistribution
from ._csv import write_csv as _write_csv
from .utils import irradiance_to_photon_flux, spectral_distribution_to_photon_distribution

Distribution = pvtrace_namespace().Distribution

# assumptions (reference solar_data.py:19-22)
water_vapor_content = 0.5  # cm
tau500 = 0.1
ozone = 0.31  # atm-cm
albedo = 0.2


def _have_pvlib() -> bool:
    try:
        import pvlib  # type: ignore  # noqa: F401

        return not getattr(pvlib, "__lscpm_shim__", False)
    except Exception:
        return False


def _is_rank_zero() -> bool:
    import os

    return int(os.environ.get("RANK", "0")) == 0


def _photon_distribution(wavelengths, spectral_irradiance, integration_time):
    return spectral_distribution_to_photon_distribution(Distribution(wavelengths, spectral_irradiance), integration_time)


def solar_data_for_place_and_time(site, tilt_angle: int, time_resolution: int = 1800, write_csv: bool = True) -> pd.DataFrame:
    """Solar position and spectral distributions for every time point of 2020 at ``site`` (reference :25-155)."""
    datetime_points = pd.date_range(start=datetime.datetime(2020, 1, 1), end=datetime.datetime(2021, 1, 1),
                                    freq=f"{time_resolution}s", tz=site.tz)
    pressure = pvlib_lite.alt2pres(site.altitude)
    sol_pos = site.get_solarposition(times=datetime_points)
    relative_airmass = site.get_airmass(times=datetime_points, solar_position=sol_pos)
    solar_data = pd.concat([sol_pos, relative_airmass], axis=1)
    solar_data = solar_data[solar_data["apparent_elevation"] > 0].copy()

    solar_data["aoi"] = pvlib_lite.irradiance.aoi(surface_tilt=tilt_angle, surface_azimuth=180,
                                                  solar_zenith=solar_data["apparent_zenith"].values,
                                                  solar_azimuth=solar_data["azimuth"].values)
    if _have_pvlib():
        from pvlib import spectrum  # type: ignore

        spec = spectrum.spectrl2(apparent_zenith=solar_data["apparent_zenith"].values, aoi=solar_data["aoi"].values,
                                 surface_tilt=tilt_angle, ground_albedo=albedo, surface_pressure=pressure,
                                 relative_airmass=solar_data["airmass_relative"].values,
                                 precipitable_water=water_vapor_content, ozone=ozone, aerosol_turbidity_500nm=tau500,
                                 dayofyear=solar_data.index.dayofyear.values)
        wavelengths = np.asarray(spec["wavelength"][11:38], dtype=float)
        poa_direct = np.asarray(spec["poa_direct"])[11:38].T
        poa_sky = np.asarray(spec["poa_sky_diffuse"])[11:38].T
    else:
        spec = pvlib_lite.spectrum.spectrl2(
            apparent_zenith=solar_data["apparent_zenith"].values, aoi=solar_data["aoi"].values, surface_tilt=tilt_angle,
            ground_albedo=albedo, surface_pressure=pressure, relative_airmass=solar_data["airmass_relative"].values,
            precipitable_water=water_vapor_content, ozone=ozone, aerosol_turbidity_500nm=tau500,
            dayofyear=solar_data.index.dayofyear.values)
        wavelengths = spec["wavelength"]            # already the [11:38] slice
        assert np.array_equal(wavelengths, SPECTRL2_SLICE_NM)
        poa_direct = spec["poa_direct"].T
        poa_sky = spec["poa_sky_diffuse"].T

    # All time points at once.  A negative ordinate (sun behind the plane) makes Distribution raise: the reference skips
    # those rows (:93-104); they keep None / NaN here and fall out with the irradiance filter below.
    poa_direct = np.asarray(poa_direct, dtype=float)
    poa_sky = np.asarray(poa_sky, dtype=float)
    ok = ~((poa_direct < 0).any(axis=1) | (poa_sky < 0).any(axis=1) | ~np.isfinite(poa_direct).all(axis=1))
    wavelengths = np.asarray(wavelengths, dtype=float)
    flux_direct = irradiance_to_photon_flux(poa_direct, wavelengths) * time_resolution     # utils.py:26-36, vectorised
    flux_sky = irradiance_to_photon_flux(poa_sky, wavelengths) * time_resolution
    direct_irr = np.where(ok, np.trapezoid(flux_direct, wavelengths, axis=1), np.nan)
    diffuse_irr = np.where(ok, np.trapezoid(flux_sky, wavelengths, axis=1), np.nan)
    direct_spectrum = np.empty(len(solar_data), dtype=object)
    diffuse_spectrum = np.empty(len(solar_data), dtype=object)
    for i in np.flatnonzero(ok):
        direct_spectrum[i] = LazyDistribution(wavelengths, flux_direct[i])
        diffuse_spectrum[i] = LazyDistribution(wavelengths, flux_sky[i])
    solar_data["direct_spectrum"] = direct_spectrum
    solar_data["diffuse_spectrum"] = diffuse_spectrum
    solar_data["direct_irradiance"] = direct_irr
    solar_data["diffuse_irradiance"] = diffuse_irr
    solar_data = solar_data[solar_data["direct_irradiance"] > 0].copy()

    if write_csv and _is_rank_zero():   # under torchrun every rank computes the table, one writes it
        _write_csv(solar_data, f"Full_data_{site.name}.csv", columns=("apparent_zenith", "zenith", "apparent_elevation", "elevation",
                                                                 "azimuth", "airmass_relative", "aoi", "direct_irradiance",
                                                                 "diffuse_irradiance"))
    logger = logging.getLogger("pvtrace").getChild("miniplant")
    logger.info(f"Generated solar data for {site.name} [lat. {site.latitude}, long. {site.longitude}] in the time range"
                f" {datetime_points.min().isoformat()} -- {datetime_points.max().isoformat()}")
    return solar_data


def solar_data_from_positions(apparent_elevation, azimuth, tilt_angle: int, index=None, spectrum=None,
                              time_resolution: int = 1800, direct_irradiance=1.0, diffuse_irradiance=1.0) -> pd.DataFrame:
    """Driver input from solar positions alone (e.g. the ``apparent_elevation`` / ``azimuth`` columns of the
    reference's committed result CSVs, the only ephemerides available offline) with one fixed spectrum."""
    from ..workloads import am15_like_photon_spectrum

    knots = am15_like_photon_spectrum() if spectrum is None else np.asarray(spectrum, dtype=float)
    dist = Distribution(knots[:, 0], knots[:, 1])
    n = len(apparent_elevation)
    frame = pd.DataFrame({"apparent_elevation": np.asarray(apparent_elevation, dtype=float),
                          "azimuth": np.asarray(azimuth, dtype=float)}, index=index)
    frame["aoi"] = pvlib_lite.irradiance.aoi(tilt_angle, 180, 90 - frame["apparent_elevation"].values, frame["azimuth"].values)
    frame["direct_spectrum"] = [dist] * n
    frame["diffuse_spectrum"] = [dist] * n
    frame["direct_irradiance"] = direct_irradiance
    frame["diffuse_irradiance"] = diffuse_irradiance
    return frame
