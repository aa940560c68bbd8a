"""Solar irradiance and spectrum incident on the reactor — mirror of reference ``miniplant/solar_data.py``.

Same entry point and output columns (``solar_data_for_place_and_time(site, tilt_angle, time_resolution)`` returns a
DataFrame with ``apparent_elevation``, ``azimuth``, ``aoi``, ``direct_spectrum`` / ``diffuse_spectrum`` photon-flux
``Distribution`` objects on the 27 wavelengths 360-690 nm and ``direct_irradiance`` / ``diffuse_irradiance``), computed
for all time points at once instead of row by row.

pvlib is not installable in the build image, so two pieces are restated here:

* **solar position** — ``compat/pvlib_lite.solar_position`` (NOAA/Meeus series + SPA refraction), within 0.015 deg of the
  ephemerides committed in the reference's result CSVs;
* **spectra** — pvlib's SPECTRL2 needs its 122-row coefficient table (extraterrestrial spectrum, water / ozone / mixed-gas
  absorption), which is not in the reference tree.  ``clear_sky_spectra`` is therefore a *SPECTRL2-like stand-in*: Bird &
  Riordan's Rayleigh and aerosol transmittances and sky-diffuse split on the same 27 wavelengths, a 5778 K black-body
  extraterrestrial spectrum, no gas absorption.  It produces plausible spectra for the drivers' workloads; it is NOT
  numerically the reference's SPECTRL2 output.  When pvlib is importable the reference's own
  calls are used instead.
"""
from __future__ import annotations

import datetime
import logging

import numpy as np
import pandas as pd

from ..compat import pvlib_lite, pvtrace_namespace
from ..workloads import SPECTRL2_SLICE_NM
from .utils import Avogadro, Planck, speed_of_light, spectral_distribution_to_photon_distribution

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


def clear_sky_spectra(apparent_zenith, aoi, surface_tilt, relative_airmass, surface_pressure, dayofyear,
                      aerosol_turbidity_500nm=tau500, ground_albedo=albedo):
    """SPECTRL2-like plane-of-array spectra on the 27-wavelength slice (W m^-2 nm^-1).

    Returns ``(poa_direct, poa_sky_diffuse)`` of shape ``[n_times, 27]``; ``poa_direct`` is negative when the sun is
    behind the plane, exactly like pvlib's (the caller drops those rows, reference solar_data.py:93-104)."""
    lam_um = SPECTRL2_SLICE_NM[None, :] * 1e-3
    lam_m = SPECTRL2_SLICE_NM[None, :] * 1e-9
    zen = np.radians(np.asarray(apparent_zenith, dtype=float))[:, None]
    cos_z = np.cos(zen)
    cos_aoi = np.cos(np.radians(np.asarray(aoi, dtype=float)))[:, None]
    am = np.asarray(relative_airmass, dtype=float)[:, None]
    am_p = am * surface_pressure / 101325.0
    doy = np.asarray(dayofyear, dtype=float)[:, None]
    # extraterrestrial spectrum: black body at 5778 K seen from 1 AU, earth-sun distance correction
    k_b = 1.380649e-23
    planck = 2 * Planck * speed_of_light ** 2 / lam_m ** 5 / np.expm1(Planck * speed_of_light / (lam_m * k_b * 5778.0))
    e0 = planck * np.pi * (6.957e8 / 1.495978707e11) ** 2 * 1e-9
    e0 = e0 * (1 + 0.034 * np.cos(2 * np.pi * doy / 365.0))
    # Bird & Riordan (1986) transmittances
    t_rayleigh = np.exp(-am_p / (lam_um ** 4 * (115.6406 - 1.335 / lam_um ** 2)))
    alpha = np.where(lam_um <= 0.5, 1.0274, 1.2060)
    tau_a = aerosol_turbidity_500nm * (lam_um / 0.5) ** -alpha
    t_aerosol = np.exp(-tau_a * am)
    direct_normal = e0 * t_rayleigh * t_aerosol
    omega = 0.945 * np.exp(-0.095 * np.log(lam_um / 0.4) ** 2)
    t_as = np.exp(-omega * tau_a * am)
    t_aa = np.exp(-(1 - omega) * tau_a * am)
    alg = np.log(1 - 0.65)
    afs = alg * (1.459 + alg * (0.1595 + alg * 0.4129))
    bfs = alg * (0.0783 + alg * (-0.3824 - alg * 0.5874))
    f_s = 1 - 0.5 * np.exp((afs + bfs * cos_z) * cos_z)
    i_rayleigh = e0 * cos_z * t_aa * (1 - t_rayleigh ** 0.95) * 0.5
    i_aerosol = e0 * cos_z * t_rayleigh ** 1.5 * t_aa * (1 - t_as) * f_s
    sky_horizontal = np.clip(i_rayleigh + i_aerosol, 0.0, None)
    tilt = np.radians(surface_tilt)
    aniso = np.clip(direct_normal / e0, 0.0, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        circumsolar = np.where(cos_z > 0.01, np.clip(cos_aoi, 0, None) / cos_z, 0.0)
    poa_sky = sky_horizontal * (aniso * circumsolar + 0.5 * (1 + np.cos(tilt)) * (1 - aniso))
    return direct_normal * cos_aoi, poa_sky


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
        wavelengths = SPECTRL2_SLICE_NM
        poa_direct, poa_sky = clear_sky_spectra(solar_data["apparent_zenith"].values, solar_data["aoi"].values, tilt_angle,
                                                solar_data["airmass_relative"].values, pressure,
                                                solar_data.index.dayofyear.values)

    direct_spectrum, diffuse_spectrum = [], []
    direct_irr = np.full(len(solar_data), np.nan)
    diffuse_irr = np.full(len(solar_data), np.nan)
    for i in range(len(solar_data)):
        # a negative ordinate (sun behind the plane) makes Distribution raise: the reference skips those rows (:93-104)
        if np.any(poa_direct[i] < 0) or np.any(poa_sky[i] < 0) or not np.all(np.isfinite(poa_direct[i])):
            direct_spectrum.append(None)
            diffuse_spectrum.append(None)
            continue
        d = _photon_distribution(wavelengths, poa_direct[i], time_resolution)
        s = _photon_distribution(wavelengths, poa_sky[i], time_resolution)
        direct_spectrum.append(d)
        diffuse_spectrum.append(s)
        direct_irr[i] = np.trapezoid(d._y, d._x)
        diffuse_irr[i] = np.trapezoid(s._y, s._x)
    solar_data["direct_spectrum"] = direct_spectrum
    solar_data["diffuse_spectrum"] = diffuse_spectrum
    solar_data["direct_irradiance"] = direct_irr
    solar_data["diffuse_irradiance"] = diffuse_irr
    solar_data = solar_data[solar_data["direct_irradiance"] > 0].copy()

    if write_csv:
        solar_data.to_csv(f"Full_data_{site.name}.csv", columns=("apparent_zenith", "zenith", "apparent_elevation", "elevation",
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
