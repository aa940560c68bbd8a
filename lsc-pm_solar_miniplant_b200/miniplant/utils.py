"""Photon-source callables and spectral unit conversion — mirror of reference ``miniplant/utils.py``.

Every class keeps the reference's name, constructor and call behaviour (they are what
``create_direct_scene`` / ``create_diffuse_scene`` bind into the light node, reference
``miniplant/scene_creator.py:295-303,321-328``).  On the CUDA path the callables are *recognised by
type* by the scene flattener (``flatten.py``) and turned into a light descriptor sampled on the
device; they remain callable on the host so that user code that emits rays by hand keeps working.
"""
from __future__ import annotations

from typing import Iterator

import numpy as np

from ..compat import pvtrace_namespace, pvlib_lite

_pv = pvtrace_namespace()
Distribution, Ray, Light, rectangular_mask = _pv.Distribution, _pv.Ray, _pv.Light, _pv.rectangular_mask
rotation_matrix, spherical_to_cart = _pv.rotation_matrix, _pv.spherical_to_cart

try:  # real pvlib when present, restated dot product otherwise (reference utils.py:7,113)
    from pvlib import irradiance  # type: ignore
except Exception:  # pragma: no cover - depends on the image
    irradiance = pvlib_lite.irradiance

# CODATA 2018 exact values (scipy.constants in the reference, utils.py:11)
Planck = 6.62607015e-34
speed_of_light = 299792458.0
Avogadro = 6.02214076e23

PLATE_HALF_SIDE = 0.47 / 2  # the light mask spans the whole 47 cm plate (reference utils.py:93)


def photon_energy(wavelength):
    """Energy in J of one photon of the given wavelength in nm (reference utils.py:14-16)."""
    return Planck * speed_of_light / (wavelength * 1e-9)


def irradiance_to_photon_flux(irradiance_per_nm, at_wavelength):
    """W m^-2 -> mol m^-2 s^-1 (reference utils.py:19-23)."""
    return irradiance_per_nm / photon_energy(at_wavelength) / Avogadro


def spectral_distribution_to_photon_distribution(distribution, integration_time=1800):
    """Spectral irradiance distribution -> photon-flux distribution in mol per integration step
    (reference utils.py:26-36).  Weighting by wavelength is what shapes the sampling CDF."""
    x = np.asarray(distribution._x, dtype=float)
    y = np.asarray(distribution._y, dtype=float)
    flux = irradiance_to_photon_flux(y, x) * integration_time
    return Distribution(distribution._x, np.asarray(flux))


class PhotonFactory:
    """Callable drawing wavelengths from a spectrum by inverse CDF (reference utils.py:39-46)."""

    def __init__(self, spectrum):
        self.spectrum = spectrum

    def __call__(self, *args, **kwargs):
        return self.spectrum.sample(np.random.uniform())


class MyLight(Light):
    """Light whose position and direction come from ONE joint generator (reference utils.py:49-76)."""

    def __init__(self, wavelength=None, position_and_direction=None, name="Light"):
        self.wavelength = wavelength
        self.position_direction = position_and_direction
        self.name = name

    def emit(self, num_rays=None) -> Iterator[Ray]:
        if not num_rays:  # None or 0: nothing is emitted (reference utils.py:58-59)
            return
        for _ in range(int(num_rays)):
            position, direction = self.position_direction()
            yield Ray(wavelength=self.wavelength(), position=position, direction=direction, source=self.name)


class VectorInverter:
    """Constant direction: the negated sun vector (reference utils.py:79-84)."""

    def __init__(self, vector):
        self.vector = vector

    def __call__(self, *args, **kwargs):
        return tuple(-component for component in self.vector)


class LightPosition:
    """Uniform point on the plate's tilted top face, which passes through the world origin
    (reference utils.py:87-99: ``inv(R_y(-tilt))`` applied to a point of the z=0 mask)."""

    def __init__(self, tilt_angle):
        self.tilt_angle = tilt_angle

    def matrix(self) -> np.ndarray:
        return np.linalg.inv(rotation_matrix(np.radians(-self.tilt_angle), (0, 1, 0)))

    def __call__(self, *args, **kwargs):
        point = np.ones(4)
        point[0:3] = rectangular_mask(PLATE_HALF_SIDE, PLATE_HALF_SIDE)
        return tuple(np.dot(self.matrix(), point)[0:3])


def create_diffuse_photon(tilt_angle: int = 30) -> np.ndarray:
    """Outward unit vector of a sky direction that sees the plate's front face: zenith ~ U(0,90) deg,
    azimuth ~ U(0,360) deg — uniform in angle, not in solid angle — redrawn until the projection on
    the plate normal is non-negative (reference utils.py:102-122)."""
    while True:
        azimuth = np.random.rand() * 360
        zenith = np.random.rand() * 90
        projection = irradiance.aoi_projection(
            surface_tilt=tilt_angle, surface_azimuth=0, solar_zenith=zenith, solar_azimuth=azimuth
        )
        if not projection < 0:
            break
    return spherical_to_cart(theta=np.deg2rad(zenith), phi=np.deg2rad(azimuth))


class IsotropicPhotonGenerator:
    """Joint (position, direction) generator for sky-diffuse light: a top-face point moved 1 m up the
    sampled sky direction, aimed back at the plate (reference utils.py:125-142)."""

    def __init__(self, tilt_angle):
        self.tilt_angle = tilt_angle
        self.base_position_generator = LightPosition(tilt_angle)

    def __call__(self, *args, **kwargs):
        on_plate = np.asarray(self.base_position_generator(), dtype=float)
        sky = create_diffuse_photon(self.tilt_angle)
        return on_plate + sky, tuple(-component for component in sky)
