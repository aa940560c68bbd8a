"""ctypes loader for the C twin of the CPU oracle (``oracle/liblscpm_oracle.so``).

TEST INFRASTRUCTURE ONLY — see the header of ``lscpm_oracle.c``.  Takes the same flattened
description structures as the CUDA library (``include/lscpm.h``) so that both consume identical input.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblscpm_oracle.so")
N_COUNTERS = 2


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "lscpm_oracle.c")
    header = os.path.join(HERE, "..", "include", "lscpm.h")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(header)):
        base = ["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-O2", "-fPIC", "-shared", "-o", LIB, src, "-lm"]
        try:
            subprocess.run(base[:1] + ["-fopenmp"] + base[1:], check=True, capture_output=True, text=True)
        except subprocess.CalledProcessError:
            subprocess.run(base, check=True, capture_output=True, text=True)
    return LIB


class COracle:
    def __init__(self, native_module, packed_scene):
        """``native_module`` = the product's ctypes structure definitions (header mirror);
        ``packed_scene`` = ``native_module.PackedScene`` holding the flattened scene."""
        self.nat = native_module
        self.scene = packed_scene
        self.lib = C.CDLL(build())
        nat = native_module
        dp, ip = nat.c_double_p, nat.c_int32_p
        self.lib.oracle_n_bins.restype = C.c_int
        self.lib.oracle_n_bins.argtypes = [C.POINTER(nat.SceneDesc)]
        self.lib.oracle_trace.restype = C.c_int
        self.lib.oracle_trace.argtypes = [C.POINTER(nat.SceneDesc), C.POINTER(nat.BatchDesc), C.POINTER(nat.SpectraDesc),
                                          C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, nat.c_int64_p]
        self.lib.oracle_emit.restype = C.c_int
        self.lib.oracle_emit.argtypes = [C.POINTER(nat.BatchDesc), C.POINTER(nat.SpectraDesc), C.c_uint64, C.c_int64,
                                         C.c_uint64, dp, dp, dp]
        self.lib.oracle_probe.restype = C.c_int
        self.lib.oracle_probe.argtypes = [C.POINTER(nat.SceneDesc), C.c_int64, dp, dp, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp]
        self.lib.oracle_follow.restype = C.c_int
        self.lib.oracle_follow.argtypes = [C.POINTER(nat.SceneDesc), C.c_int64, dp, dp, dp, C.c_uint32, C.c_uint64,
                                           C.c_uint64, C.c_int32, ip, ip, ip, dp, dp, dp, ip]
        self.n_bins = self.lib.oracle_n_bins(C.byref(self.scene.desc))
        if self.n_bins < 0:
            raise RuntimeError(f"oracle_n_bins failed: {self.n_bins}")

    def _spectra(self, spectra):
        packed = self.nat.PackedSpectra(list(spectra) if spectra else [])
        return packed

    def trace(self, batch, spectra=None, photon_begin=0, photon_count=1000, seed=0, threads=0) -> np.ndarray:
        out = np.zeros(self.n_bins + N_COUNTERS, dtype=np.int64)
        sp = self._spectra(spectra)
        rc = self.lib.oracle_trace(C.byref(self.scene.desc), C.byref(batch), C.byref(sp.desc), photon_begin,
                                   photon_count, seed, threads, out.ctypes.data_as(self.nat.c_int64_p))
        if rc:
            raise RuntimeError(f"oracle_trace failed: {rc}")
        return out

    def emit(self, batch, n, spectra=None, photon_begin=0, seed=0):
        pos, direc, wl = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
        sp = self._spectra(spectra)
        dp = self.nat.c_double_p
        rc = self.lib.oracle_emit(C.byref(batch), C.byref(sp.desc), photon_begin, n, seed,
                                  pos.ctypes.data_as(dp), direc.ctypes.data_as(dp), wl.ctypes.data_as(dp))
        if rc:
            raise RuntimeError(f"oracle_emit failed: {rc}")
        return pos, direc, wl

    def probe(self, pos, direc) -> dict:
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        direc = np.ascontiguousarray(direc, dtype=np.float64).reshape(-1, 3)
        n = len(pos)
        out = dict(hit=np.zeros(n, np.int32), container=np.zeros(n, np.int32), adjacent=np.zeros(n, np.int32),
                   face=np.zeros(n, np.int32), distance=np.zeros(n), point=np.zeros((n, 3)), normal=np.zeros((n, 3)),
                   reflectivity=np.zeros(n), reflected=np.zeros((n, 3)), refracted=np.zeros((n, 3)))
        dp, ip = self.nat.c_double_p, self.nat.c_int32_p
        rc = self.lib.oracle_probe(C.byref(self.scene.desc), n, pos.ctypes.data_as(dp), direc.ctypes.data_as(dp),
                                   out["hit"].ctypes.data_as(ip), out["container"].ctypes.data_as(ip),
                                   out["adjacent"].ctypes.data_as(ip), out["face"].ctypes.data_as(ip),
                                   out["distance"].ctypes.data_as(dp), out["point"].ctypes.data_as(dp),
                                   out["normal"].ctypes.data_as(dp), out["reflectivity"].ctypes.data_as(dp),
                                   out["reflected"].ctypes.data_as(dp), out["refracted"].ctypes.data_as(dp))
        if rc:
            raise RuntimeError(f"oracle_probe failed: {rc}")
        return out

    def follow(self, pos, direc, wavelength, batch_id=0, photon_begin=0, seed=0, max_steps=64) -> dict:
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        direc = np.ascontiguousarray(direc, dtype=np.float64).reshape(-1, 3)
        wavelength = np.ascontiguousarray(wavelength, dtype=np.float64).reshape(-1)
        n, m = len(pos), max(1, max_steps)
        out = dict(n_steps=np.zeros(n, np.int32), event=np.full((n, m), -1, np.int32), node=np.full((n, m), -1, np.int32),
                   pos=np.zeros((n, m, 3)), dir=np.zeros((n, m, 3)), wavelength=np.zeros((n, m)),
                   final_bin=np.zeros(n, np.int32))
        dp, ip = self.nat.c_double_p, self.nat.c_int32_p
        rc = self.lib.oracle_follow(C.byref(self.scene.desc), n, pos.ctypes.data_as(dp), direc.ctypes.data_as(dp),
                                    wavelength.ctypes.data_as(dp), batch_id, photon_begin, seed, max_steps,
                                    out["n_steps"].ctypes.data_as(ip), out["event"].ctypes.data_as(ip),
                                    out["node"].ctypes.data_as(ip), out["pos"].ctypes.data_as(dp),
                                    out["dir"].ctypes.data_as(dp), out["wavelength"].ctypes.data_as(dp),
                                    out["final_bin"].ctypes.data_as(ip))
        if rc:
            raise RuntimeError(f"oracle_follow failed: {rc}")
        return out
