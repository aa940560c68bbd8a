"""Shared helpers for the parity tests (test code only)."""
from __future__ import annotations

import numpy as np

Z_999 = 3.2905  # two-sided 99.9 % normal quantile


def standard_scene(tilt=0.0, elevation=90.0, azimuth=180.0, wavelength=585.0, include_dye=True, spectrum=None,
                   diffuse=False):
    from lscpm_b200.miniplant.scene_creator import create_direct_scene, create_diffuse_scene
    from lscpm_b200.miniplant.utils import PhotonFactory

    if spectrum is not None:
        from lscpm_b200.compat.pvtrace_api import Distribution

        fn = PhotonFactory(Distribution(spectrum[:, 0], spectrum[:, 1]))
    else:
        fn = ConstantWavelength(wavelength)
    if diffuse:
        return create_diffuse_scene(tilt_angle=tilt, solar_spectrum_function=fn, include_dye=include_dye)
    return create_direct_scene(tilt_angle=tilt, solar_elevation=elevation, solar_azimuth=azimuth,
                               solar_spectrum_function=fn, include_dye=include_dye)


class ConstantWavelength:
    def __init__(self, nm):
        self.nm = float(nm)

    def __call__(self):
        return self.nm


# Synthetic "AM1.5G-like" spectrum on the SPECTRL2 grid slice [11:38] the reference samples
# (miniplant/solar_data.py:92-101); SURVEY.md section 8(d) config 2.  W m^-2 nm^-1, smooth and positive.
SPECTRL2_SLICE_NM = np.array([360, 370, 380, 390, 400, 410, 420, 430, 440, 450, 460, 470, 480, 490, 500, 510, 520,
                              530, 540, 550, 570, 593, 610, 630, 656, 667.6, 690], dtype=float)
AM15_LIKE_W = np.array([0.55, 0.68, 0.66, 0.72, 1.02, 1.12, 1.12, 1.07, 1.26, 1.42, 1.44, 1.42, 1.47, 1.40, 1.42,
                        1.42, 1.37, 1.42, 1.38, 1.40, 1.36, 1.33, 1.35, 1.30, 1.22, 1.30, 1.12], dtype=float)


def am15_like_photon_spectrum() -> np.ndarray:
    """(27, 2) knots of the photon-flux distribution: irradiance weighted by wavelength, as
    reference miniplant/utils.py:14-36 does (constant factors drop out of the normalised CDF)."""
    return np.stack([SPECTRL2_SLICE_NM, AM15_LIKE_W * SPECTRL2_SLICE_NM], axis=1)


def assert_counts_agree(names, a, b, z=Z_999, label=""):
    """Two independent multinomial samples of equal size: every bin's difference inside the 99.9 % band."""
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    assert a.sum() == b.sum(), (a.sum(), b.sum())
    bad = []
    for name, ca, cb in zip(names, a, b):
        if ca == 0 and cb == 0:
            continue
        allowed = z * np.sqrt(max(ca + cb, 4))
        if abs(int(ca) - int(cb)) > allowed:
            bad.append((name, int(ca), int(cb), round(float(allowed), 1)))
    assert not bad, f"{label}: bins outside the 99.9% band: {bad}"


def grouped(names, counts) -> dict:
    """Pool the per-node bins into the families the reference talks about."""
    groups = {}
    for name, c in zip(names, counts):
        parts = name.split("/")
        if parts[0] == "react":
            key = "react"
        elif parts[0] == "absorb":
            key = "absorb/" + ("capillary" if parts[1].startswith("Capillary") else
                               "mixture" if parts[1].startswith("Reaction") else f"plate/{parts[2]}")
        elif parts[0] == "exit" and len(parts) == 3:
            key = "exit/" + ("edge" if parts[2] in ("-x", "+x", "-y", "+y") else parts[2])
        else:
            key = name
        groups[key] = groups.get(key, 0) + int(c)
    return groups
