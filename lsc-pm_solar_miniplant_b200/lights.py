"""Light descriptors built directly from the drivers' per-row parameters, without constructing a scene graph.

The drivers call ``run_direct_simulation`` / ``run_diffuse_simulation`` once per time point (reference
``miniplant/full_simulation.py:77-97``, ``angle_optimization_simulation.py:73-81``), each of which rebuilds the whole
scene only to change the light.  These helpers restate the light half of ``create_direct_scene`` /
``create_diffuse_scene`` (reference ``miniplant/scene_creator.py:288-304,321-328`` with ``utils.py:87-142``) as
``LightArrays``; ``tests/test_solar_stage.py`` checks them against flattening the real scenes.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .flatten import LIGHT_DIFFUSE, LIGHT_DIRECT, LightArrays

PLATE_HALF_SIDE = 0.47 / 2


def _mask_pose(tilt_angle: float) -> np.ndarray:
    """``inv(R_y(-tilt))`` = rotation by +tilt about y (reference utils.py:94): mask plane -> world, 3x4."""
    t = np.radians(tilt_angle)
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0.0, s, 0.0], [0.0, 1.0, 0.0, 0.0], [-s, 0.0, c, 0.0]])


def _spectrum(spectrum) -> Optional[np.ndarray]:
    if spectrum is None:
        return None
    if hasattr(spectrum, "_x"):
        return np.stack([np.asarray(spectrum._x, float), np.asarray(spectrum._y, float)], axis=1)
    return np.asarray(spectrum, dtype=float).reshape(-1, 2)


def direct_light(tilt_angle: float, solar_elevation: float, solar_azimuth: float, spectrum=None,
                 wavelength_nm: float = 555.0) -> LightArrays:
    theta, phi = np.deg2rad(-solar_elevation + 90), np.deg2rad(-solar_azimuth + 180)
    sun = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
    return LightArrays(kind=LIGHT_DIRECT, mask_half=np.array([PLATE_HALF_SIDE, PLATE_HALF_SIDE]), mask_pose=_mask_pose(tilt_angle),
                       direction=-sun, back_off=1.0, tilt_deg=float(tilt_angle), wavelength_nm=float(wavelength_nm),
                       spectrum=_spectrum(spectrum))


def diffuse_light(tilt_angle: float, spectrum=None, wavelength_nm: float = 555.0) -> LightArrays:
    return LightArrays(kind=LIGHT_DIFFUSE, mask_half=np.array([PLATE_HALF_SIDE, PLATE_HALF_SIDE]), mask_pose=_mask_pose(tilt_angle),
                       direction=np.zeros(3), back_off=1.0, tilt_deg=float(tilt_angle), wavelength_nm=float(wavelength_nm),
                       spectrum=_spectrum(spectrum))
