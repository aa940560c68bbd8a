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


class Location:
    def __init__(self, latitude, longitude, tz="UTC", altitude=0, name=None):
        self.latitude = latitude
        self.longitude = longitude
        self.tz = tz
        self.altitude = altitude
        self.name = name

    def __repr__(self):
        return f"Location(name={self.name!r}, latitude={self.latitude}, longitude={self.longitude}, tz={self.tz!r})"
