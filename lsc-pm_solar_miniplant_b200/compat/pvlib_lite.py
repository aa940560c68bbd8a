"""The two pvlib names the hot path touches, restated (pvlib==0.9.5 is not installable here).

* ``irradiance.aoi_projection`` is called once per trial direction inside the diffuse photon
  generator (reference ``miniplant/utils.py:113``).  It is the dot product of the panel normal and
  the sun vector: ``cos(tilt)cos(zen) + sin(tilt)sin(zen)cos(az - surface_az)`` (degrees in).
* ``location.Location`` is the site record used by ``miniplant/locations.py:3-32`` and the
  drivers; only the attributes are kept (solar-position / SPECTRL2 methods are CPU input
  preparation).
"""
from __future__ import annotations

import numpy as np


class irradiance:  # namespace mirroring ``pvlib.irradiance``
    @staticmethod
    def aoi_projection(surface_tilt, surface_azimuth, solar_zenith, solar_azimuth):
        tilt = np.radians(surface_tilt)
        zen = np.radians(solar_zenith)
        daz = np.radians(np.asarray(solar_azimuth, dtype=float) - surface_azimuth)
        proj = np.cos(tilt) * np.cos(zen) + np.sin(tilt) * np.sin(zen) * np.cos(daz)
        return np.clip(proj, -1.0, 1.0)

    @staticmethod
    def aoi(surface_tilt, surface_azimuth, solar_zenith, solar_azimuth):
        return np.degrees(np.arccos(irradiance.aoi_projection(
            surface_tilt, surface_azimuth, solar_zenith, solar_azimuth)))


def alt2pres(altitude):
    """Standard-atmosphere pressure in Pa at an altitude in m (pvlib.atmosphere.alt2pres)."""
    return 100.0 * ((44331.514 - altitude) / 11880.516) ** (1 / 0.1902632)


def solar_position(times, latitude, longitude, altitude=0.0, pressure=None, temperature=12.0):
    """Sun position for a tz-aware ``DatetimeIndex`` (NOAA / Meeus low-precision series, ~0.01 deg) with the SPA
    refraction correction pvlib applies for the apparent elevation.  Stands in for
    ``Location.get_solarposition`` (reference miniplant/solar_data.py:48); checked against the ephemerides committed
    in the reference's result CSVs (tests/test_solar_stage.py).  Returns a DataFrame with pvlib's column names."""
    import pandas as pd

    t = pd.DatetimeIndex(times)
    utc = t.tz_convert("UTC") if t.tz is not None else t
    jd = utc.to_julian_date().values
    T = (jd - 2451545.0) / 36525.0
    L0 = (280.46646 + T * (36000.76983 + 0.0003032 * T)) % 360
    M = np.radians(357.52911 + T * (35999.05029 - 0.0001537 * T))
    e = 0.016708634 - T * (0.000042037 + 0.0000001267 * T)
    C = (np.sin(M) * (1.914602 - T * (0.004817 + 0.000014 * T)) + np.sin(2 * M) * (0.019993 - 0.000101 * T)
         + np.sin(3 * M) * 0.000289)
    omega = np.radians(125.04 - 1934.136 * T)
    lam = np.radians(L0 + C - 0.00569 - 0.00478 * np.sin(omega))
    eps = np.radians(23 + (26 + (21.448 - T * (46.815 + T * (0.00059 - T * 0.001813))) / 60) / 60 + 0.00256 * np.cos(omega))
    decl = np.arcsin(np.sin(eps) * np.sin(lam))
    y = np.tan(eps / 2) ** 2
    L0r = np.radians(L0)
    eot = 4 * np.degrees(y * np.sin(2 * L0r) - 2 * e * np.sin(M) + 4 * e * y * np.sin(M) * np.cos(2 * L0r)
                         - 0.5 * y * y * np.sin(4 * L0r) - 1.25 * e * e * np.sin(2 * M))
    minutes = (utc.hour * 60 + utc.minute + utc.second / 60.0).values
    hour_angle = np.radians(((minutes + eot + 4 * longitude) % 1440) / 4 - 180)
    lat = np.radians(latitude)
    cosz = np.sin(lat) * np.sin(decl) + np.cos(lat) * np.cos(decl) * np.cos(hour_angle)
    elevation = 90 - np.degrees(np.arccos(np.clip(cosz, -1, 1)))
    azimuth = (np.degrees(np.arctan2(np.sin(hour_angle), np.cos(hour_angle) * np.sin(lat) - np.tan(decl) * np.cos(lat))) + 180) % 360
    if pressure is None:
        pressure = alt2pres(altitude)
    with np.errstate(divide="ignore", invalid="ignore"):
        refraction = (pressure / 100 / 1010.0) * (283.0 / (273 + temperature)) * 1.02 / (
            60 * np.tan(np.radians(elevation + 10.3 / (elevation + 5.11))))
    refraction = np.where(elevation >= -(0.26667 + 0.5667), refraction, 0.0)
    apparent = elevation + refraction
    return pd.DataFrame({"apparent_zenith": 90 - apparent, "zenith": 90 - elevation, "apparent_elevation": apparent,
                         "elevation": elevation, "azimuth": azimuth, "equation_of_time": eot}, index=t)


def relative_airmass(apparent_zenith):
    """Kasten & Young (1989), pvlib's default model for ``Location.get_airmass``."""
    z = np.asarray(apparent_zenith, dtype=float)
    with np.errstate(invalid="ignore"):
        am = 1.0 / (np.cos(np.radians(z)) + 0.50572 * (96.07995 - z) ** -1.6364)
    return np.where(z > 90, np.nan, am)


class Location:
    """Site record with the two ``pvlib.location.Location`` methods the solar stage calls."""

    def __init__(self, latitude, longitude, tz="UTC", altitude=0, name=None):
        self.latitude = latitude
        self.longitude = longitude
        self.tz = tz
        self.altitude = altitude
        self.name = name

    def get_solarposition(self, times, pressure=None, temperature=12.0):
        return solar_position(times, self.latitude, self.longitude, self.altitude, pressure, temperature)

    def get_airmass(self, times=None, solar_position=None):
        import pandas as pd

        if solar_position is None:
            solar_position = self.get_solarposition(times)
        rel = relative_airmass(solar_position["apparent_zenith"].values)
        return pd.DataFrame({"airmass_relative": rel, "airmass_absolute": rel * alt2pres(self.altitude) / 101325.0},
                            index=solar_position.index)

    def __repr__(self):
        return f"Location(name={self.name!r}, latitude={self.latitude}, longitude={self.longitude}, tz={self.tz!r})"
