"""The two pvlib names the hot path touches, restated (pvlib==0.9.5 is not installable here).

* ``irradiance.aoi_projection`` is called once per trial direction inside the diffuse photon
  generator (reference ``miniplant/utils.py:113``).  It is the dot product of the panel normal and
  the sun vector: ``cos(tilt)cos(zen) + sin(tilt)sin(zen)cos(az - surface_az)`` (degrees in).
* ``location.Location`` is the site record used by ``miniplant/locations.py:3-32`` and the
  drivers, with the two methods the solar stage calls (``get_solarposition``, ``get_airmass``).
* ``spectrum.spectrl2`` (reference ``miniplant/solar_data.py:69-80``) is restated for the 27 wavelengths
  360-690 nm the reference keeps (``[11:38]``, ``solar_data.py:94-101``): Bird & Riordan's published SPCTRAL2
  model (SERI/TR-215-2436, 1986) with the rows of its Table 1 that fall in that range.  Checked against the
  reference's own output: the irradiances recoverable from its committed result CSVs
  (``direct_reacted / simulation_direct / area``) are reproduced to 2e-5 relative for both the direct and the
  sky-diffuse component over all ~80 000 time points of the four sites (tests/test_solar_stage.py,
  tests/golden/reference_irradiances.csv).
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


# Bird & Riordan (1986) Table 1, rows 11..37 of pvlib's spectrl2 coefficient table (the slice the reference keeps):
# wavelength [um], extraterrestrial spectral irradiance [W m^-2 um^-1], water-vapour, ozone and uniformly-mixed-gas
# absorption coefficients
_SPECTRL2_UVVIS = np.array([
    [0.360, 975.9, 0, 0, 0], [0.370, 1119.9, 0, 0, 0], [0.380, 1103.8, 0, 0, 0], [0.390, 1033.8, 0, 0, 0],
    [0.400, 1479.1, 0, 0, 0], [0.410, 1701.3, 0, 0, 0], [0.420, 1740.4, 0, 0, 0], [0.430, 1587.2, 0, 0, 0],
    [0.440, 1837.0, 0, 0, 0], [0.450, 2005.0, 0, 0.003, 0], [0.460, 2043.0, 0, 0.006, 0], [0.470, 1987.0, 0, 0.009, 0],
    [0.480, 2027.0, 0, 0.014, 0], [0.490, 1896.0, 0, 0.021, 0], [0.500, 1909.0, 0, 0.030, 0], [0.510, 1927.0, 0, 0.040, 0],
    [0.520, 1831.0, 0, 0.048, 0], [0.530, 1891.0, 0, 0.063, 0], [0.540, 1898.0, 0, 0.075, 0], [0.550, 1892.0, 0, 0.085, 0],
    [0.570, 1840.0, 0, 0.120, 0], [0.593, 1768.0, 0.075, 0.119, 0], [0.610, 1728.0, 0, 0.120, 0], [0.630, 1658.0, 0, 0.090, 0],
    [0.656, 1524.0, 0, 0.065, 0], [0.6676, 1531.0, 0, 0.051, 0], [0.690, 1420.0, 0.016, 0.028, 0.15]])
SPECTRL2_UVVIS_SLICE = slice(11, 38)   # where these rows sit in pvlib's 122-row table


def _spectrl2_transmittances(lam, cos_z, airmass_rel, pressure, water, ozone_cm, tau_a, omega):
    """Rayleigh, aerosol, water-vapour, ozone, mixed-gas, aerosol-scattering and aerosol-absorption transmittances
    (Bird & Riordan eqs. 2-4 ... 2-11, 3-6, 3-7) at relative air mass ``airmass_rel``."""
    a_w, a_o, a_u = (_SPECTRL2_UVVIS[:, i][:, None] for i in (2, 3, 4))
    airmass = airmass_rel * pressure / 101300.0
    rayleigh = np.exp(-airmass / (lam ** 4 * (115.6406 - 1.335 / lam ** 2)))
    aerosol = np.exp(-tau_a * airmass_rel)
    aw_m = a_w * water * airmass_rel
    vapour = np.exp(-0.2385 * aw_m / (1 + 20.07 * aw_m) ** 0.45)
    h0 = 22.0 / 6370.0                                           # ozone layer at 22 km
    ozone_mass = (1 + h0) / np.sqrt(cos_z ** 2 + 2 * h0)
    ozone = np.exp(-a_o * ozone_cm * ozone_mass)
    au_m = a_u * airmass
    mixed = np.exp(-1.41 * au_m / (1 + 118.3 * au_m) ** 0.45)
    scattering = np.exp(-omega * tau_a * airmass_rel)
    absorption = np.exp(-(1 - omega) * tau_a * airmass_rel)
    return rayleigh, aerosol, vapour, ozone, mixed, scattering, absorption


class spectrum:  # namespace mirroring ``pvlib.spectrum``
    @staticmethod
    def spectrl2(apparent_zenith, aoi, surface_tilt, ground_albedo, surface_pressure, relative_airmass,
                 precipitable_water, ozone, aerosol_turbidity_500nm, dayofyear, scattering_albedo_400nm=0.945,
                 alpha=1.14, wavelength_variation_factor=0.095, aerosol_asymmetry_factor=0.65):
        """Bird & Riordan clear-sky spectra on a tilted plane, pvlib's argument names and result keys, vectorised over
        time points; arrays are ``[27, n_times]`` — the ``[11:38]`` rows of pvlib's output — in W m^-2 nm^-1."""
        lam = _SPECTRL2_UVVIS[:, 0][:, None]
        h0_spec = _SPECTRL2_UVVIS[:, 1][:, None] / 1000.0
        zen = np.radians(np.atleast_1d(np.asarray(apparent_zenith, dtype=float)))[None, :]
        cos_z = np.cos(zen)
        cos_aoi = np.cos(np.radians(np.atleast_1d(np.asarray(aoi, dtype=float))))[None, :]
        am = np.atleast_1d(np.asarray(relative_airmass, dtype=float))[None, :]
        day_angle = 2 * np.pi * (np.atleast_1d(np.asarray(dayofyear, dtype=float))[None, :] - 1) / 365.0
        distance = (1.00011 + 0.034221 * np.cos(day_angle) + 0.00128 * np.sin(day_angle)
                    + 0.000719 * np.cos(2 * day_angle) + 0.000077 * np.sin(2 * day_angle))
        extra = h0_spec * distance
        tau_a = aerosol_turbidity_500nm * (lam / 0.5) ** -alpha
        omega = scattering_albedo_400nm * np.exp(-wavelength_variation_factor * np.log(lam / 0.4) ** 2)
        t_r, t_a, t_w, t_o, t_u, t_as, t_aa = _spectrl2_transmittances(
            lam, cos_z, am, surface_pressure, precipitable_water, ozone, tau_a, omega)
        dni = extra * t_r * t_a * t_w * t_o * t_u
        alg = np.log(1 - aerosol_asymmetry_factor)
        afs = alg * (1.459 + alg * (0.1595 + alg * 0.4129))
        bfs = alg * (0.0783 + alg * (-0.3824 - alg * 0.5874))
        f_s = 1 - 0.5 * np.exp((afs + bfs * cos_z) * cos_z)
        f_s_prime = 1 - 0.5 * np.exp((afs + bfs / 1.8) / 1.8)
        c_s = np.where(lam <= 0.45, (lam + 0.55) ** 1.8, 1.0)
        common = extra * cos_z * t_o * t_u * t_w * t_aa
        rayleigh_diffuse = common * (1 - t_r ** 0.95) * 0.5
        aerosol_diffuse = common * t_r ** 1.5 * (1 - t_as) * f_s
        # sky reflectivity from the transmittances at air mass 1.8 (mixed gas, water vapour, aerosol absorption: the
        # combination that reproduces the reference's sky-diffuse irradiances to 2e-5; with ozone in place of the
        # mixed gas they come out 0.25 % low)
        p_r, _, p_w, _, p_u, p_as, p_aa = _spectrl2_transmittances(
            lam, cos_z, np.full_like(am, 1.8), surface_pressure, precipitable_water, ozone, tau_a, omega)
        sky_reflectivity = p_u * p_w * p_aa * (0.5 * (1 - p_r) + (1 - f_s_prime) * p_r * (1 - p_as))
        ground_diffuse = ((dni * cos_z + rayleigh_diffuse + aerosol_diffuse) * sky_reflectivity * ground_albedo
                          / (1 - sky_reflectivity * ground_albedo))
        dhi = (rayleigh_diffuse + aerosol_diffuse + ground_diffuse) * c_s
        ghi = dni * cos_z + dhi
        tilt = np.radians(surface_tilt)
        ratio = dni / extra
        with np.errstate(divide="ignore", invalid="ignore"):
            poa_sky = dhi * (ratio * np.clip(cos_aoi, 0, None) / cos_z + 0.5 * (1 + np.cos(tilt)) * (1 - ratio))
        poa_ground = 0.5 * ghi * ground_albedo * (1 - np.cos(tilt))
        poa_direct = dni * cos_aoi       # negative with the sun behind the plane, like pvlib's (solar_data.py:102-104)
        return {"wavelength": _SPECTRL2_UVVIS[:, 0] * 1000.0, "dni_extra": extra, "dhi": dhi, "dni": dni,
                "poa_sky_diffuse": poa_sky, "poa_ground_diffuse": poa_ground, "poa_direct": poa_direct,
                "poa_global": poa_direct + poa_sky + poa_ground}


class _Zone:
    """The one ``pytz`` timezone method the reference calls on ``Location.pytz`` (``localize``,
    experimental_comparison/simulate_experimental_conditions.py:15-16), on top of ``zoneinfo``."""

    def __init__(self, name):
        from zoneinfo import ZoneInfo

        self.zone = name
        self._info = ZoneInfo(name)

    def localize(self, naive):
        return naive.replace(tzinfo=self._info)

    def __repr__(self):
        return f"<tz {self.zone}>"


class Location:
    """Site record with the two ``pvlib.location.Location`` methods the solar stage calls."""

    def __init__(self, latitude, longitude, tz="UTC", altitude=0, name=None):
        self.latitude = latitude
        self.longitude = longitude
        self.tz = tz
        self.altitude = altitude
        self.name = name
        try:
            import pytz  # type: ignore

            self.pytz = pytz.timezone(tz)
        except ImportError:
            self.pytz = _Zone(tz)

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
