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

from . import _native as nat
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


# --------------------------------------------------------------------------------------
# column-wise descriptors for the drivers (thousands of time points per launch)
# --------------------------------------------------------------------------------------


class BatchTable:
    """``LscpmBatchDesc[n]`` as a structured NumPy array (``_native.BATCH_DTYPE``) plus the spectra its rows refer to:
    ``spectra`` is ``None`` (constant wavelengths), a list of (N_i, 2) knot arrays, or one (M, N, 2) array.  Built
    column-wise — no per-row Python objects; ``backend.Plan`` / ``backend.trace_lights`` take it as it is."""

    def __init__(self, desc: np.ndarray, spectra=None):
        assert desc.dtype == nat.BATCH_DTYPE
        self.desc = np.ascontiguousarray(desc)
        self.spectra = spectra

    def __len__(self) -> int:
        return len(self.desc)

    def light(self, i: int) -> LightArrays:
        """Row ``i`` as a ``LightArrays`` (what a per-row tracer — the CPU oracle in the tests — consumes)."""
        row = self.desc[i]
        spec = int(row["spectrum"])
        return LightArrays(kind=int(row["light_kind"]), mask_half=np.array(row["mask_half"]), mask_pose=np.array(row["mask_pose"]).reshape(3, 4),
                           direction=np.array(row["direction"]), back_off=float(row["back_off"]), tilt_deg=float(row["tilt_deg"]),
                           wavelength_nm=float(row["wavelength_nm"]),
                           spectrum=None if spec < 0 else np.asarray(self.spectra[spec], dtype=np.float64))

    def __getitem__(self, i: int) -> LightArrays:
        return self.light(i)

    def __iter__(self):
        return (self.light(i) for i in range(len(self)))

    def take(self, rows) -> "BatchTable":
        """Sub-table (same spectra; the rows keep their batch ids and spectrum indices)."""
        return BatchTable(self.desc[np.asarray(rows, dtype=np.int64)], self.spectra)

    @staticmethod
    def concatenate(tables) -> "BatchTable":
        """Stack tables; (M, N, 2) spectra blocks are stacked as well and the spectrum indices shifted."""
        tables = list(tables)
        descs, blocks, offset = [], [], 0
        for t in tables:
            d = t.desc.copy()
            if t.spectra is not None:
                block = np.asarray(t.spectra, dtype=np.float64)
                if block.ndim != 3:
                    raise ValueError("only (M, N, 2) spectra blocks can be concatenated")
                d["spectrum"] = np.where(d["spectrum"] >= 0, d["spectrum"] + offset, -1)
                blocks.append(block)
                offset += len(block)
            descs.append(d)
        desc = np.concatenate(descs)
        desc["batch_id"] = np.arange(len(desc), dtype=np.uint32)
        return BatchTable(desc, np.concatenate(blocks, axis=0) if blocks else None)


def _table(n: int, kind: int, tilt_angle: float, wavelength_nm: float, spectrum_index, first_batch_id: int) -> np.ndarray:
    desc = np.zeros(n, dtype=nat.BATCH_DTYPE)
    desc["light_kind"] = kind
    desc["spectrum"] = -1 if spectrum_index is None else np.asarray(spectrum_index, dtype=np.int32)
    desc["batch_id"] = np.arange(first_batch_id, first_batch_id + n, dtype=np.uint32)
    desc["wavelength_nm"] = wavelength_nm
    desc["mask_half"] = PLATE_HALF_SIDE
    desc["mask_pose"] = _mask_pose(tilt_angle).reshape(12)
    desc["back_off"] = 1.0
    desc["tilt_deg"] = float(tilt_angle)
    return desc


def dedupe_spectra(flux: np.ndarray, x: np.ndarray):
    """(n, N) ordinates on the common abscissae ``x`` -> ((M, N, 2) distinct spectra, index[n])."""
    flux = np.ascontiguousarray(flux, dtype=np.float64)
    keys = flux.view(np.dtype((np.void, flux.dtype.itemsize * flux.shape[1]))).reshape(-1)
    _, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
    distinct = flux[first]
    block = np.empty((len(distinct), flux.shape[1], 2), dtype=np.float64)
    block[:, :, 0] = np.asarray(x, dtype=np.float64)
    block[:, :, 1] = distinct
    return block, inverse.astype(np.int32).reshape(-1)


def direct_batches(tilt_angle: float, solar_elevation, solar_azimuth, flux=None, x=None, wavelength_nm: float = 555.0,
                   first_batch_id: int = 0) -> BatchTable:
    """One direct-beam batch per (elevation, azimuth) pair — ``direct_light`` for whole columns.  ``flux`` (n, N) are
    the ordinates of each row's sampling distribution on the abscissae ``x`` (N,)."""
    elevation = np.asarray(solar_elevation, dtype=np.float64).reshape(-1)
    azimuth = np.asarray(solar_azimuth, dtype=np.float64).reshape(-1)
    theta, phi = np.deg2rad(-elevation + 90), np.deg2rad(-azimuth + 180)
    sun = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=1)
    spectra = index = None
    if flux is not None:
        spectra, index = dedupe_spectra(flux, x)
    desc = _table(len(elevation), LIGHT_DIRECT, tilt_angle, wavelength_nm, index, first_batch_id)
    desc["direction"] = -sun
    return BatchTable(desc, spectra)


def diffuse_batches(tilt_angle: float, n: int = None, flux=None, x=None, wavelength_nm: float = 555.0,
                    first_batch_id: int = 0) -> BatchTable:
    """One sky-diffuse batch per row of ``flux`` (or ``n`` constant-wavelength batches)."""
    spectra = index = None
    if flux is not None:
        spectra, index = dedupe_spectra(flux, x)
        n = len(index)
    return BatchTable(_table(int(n), LIGHT_DIFFUSE, tilt_angle, wavelength_nm, index, first_batch_id), spectra)
