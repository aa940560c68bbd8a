"""Named workloads of BASELINE.json (synthetic inputs shared by bench.py, the tests and the oracle).

The reference has no AM1.5G table (its spectra come from pvlib's SPECTRL2 at run time, reference
``miniplant/solar_data.py:69-101``) and pvlib is not installable here, so config 2's spectrum is a fixed
smooth positive table on the very 27 wavelengths the reference samples (the ``[11:38]`` slice of the
SPECTRL2 grid).  Parity only needs both sides to consume the same table.
"""
from __future__ import annotations

import numpy as np

SPECTRL2_SLICE_NM = np.array([360, 370, 380, 390, 400, 410, 420, 430, 440, 450, 460, 470, 480, 490, 500, 510, 520,
                              530, 540, 550, 570, 593, 610, 630, 656, 667.6, 690], dtype=float)
AM15_LIKE_W = np.array([0.55, 0.68, 0.66, 0.72, 1.02, 1.12, 1.12, 1.07, 1.26, 1.42, 1.44, 1.42, 1.47, 1.40, 1.42,
                        1.42, 1.37, 1.42, 1.38, 1.40, 1.36, 1.33, 1.35, 1.30, 1.22, 1.30, 1.12], dtype=float)


def am15_like_photon_spectrum() -> np.ndarray:
    """(27, 2) knots of the photon-flux distribution: irradiance weighted by wavelength, as reference
    ``miniplant/utils.py:14-36`` does (constant factors drop out of the normalised CDF)."""
    return np.stack([SPECTRL2_SLICE_NM, AM15_LIKE_W * SPECTRL2_SLICE_NM], axis=1)


class ConstantWavelength:
    """Picklable constant-wavelength callable (the reference uses ``lambda: 555``)."""

    def __init__(self, nm):
        self.nm = float(nm)

    def __call__(self):
        return self.nm


def standard_scene(tilt=0.0, elevation=90.0, azimuth=180.0, wavelength=585.0, include_dye=True, spectrum=None,
                   diffuse=False, **kwargs):
    """The reference's scene for one light configuration (``create_direct_scene`` / ``create_diffuse_scene``)."""
    from .compat.pvtrace_api import Distribution
    from .miniplant.scene_creator import create_diffuse_scene, create_direct_scene
    from .miniplant.utils import PhotonFactory

    fn = PhotonFactory(Distribution(spectrum[:, 0], spectrum[:, 1])) if spectrum is not None else ConstantWavelength(wavelength)
    if diffuse:
        return create_diffuse_scene(tilt_angle=tilt, solar_spectrum_function=fn, include_dye=include_dye)
    return create_direct_scene(tilt_angle=tilt, solar_elevation=elevation, solar_azimuth=azimuth,
                               solar_spectrum_function=fn, include_dye=include_dye, **kwargs)
