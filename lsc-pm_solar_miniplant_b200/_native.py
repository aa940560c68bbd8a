"""ctypes binding of the C-ABI in ``include/lscpm.h`` (``csrc/liblscpm.so``).

The structures mirror the header field by field.  Loading fails loudly: there is no Python or CPU
fallback for the traced path — if the CUDA library has not been built (``python -c "import
__graft_entry__ as g; g.build()"``) every product entry point raises ``LscpmLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

LIB_NAME = "liblscpm.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", LIB_NAME)
N_COUNTERS = 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class LscpmLibraryError(RuntimeError):
    pass


class LscpmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lscpm error {code}: {message}")
        self.code = code


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("node_geom", c_int32_p),
        ("node_parent", c_int32_p),
        ("node_material", c_int32_p),
        ("node_size", c_double_p),
        ("node_pose", c_double_p),
        ("node_name", C.POINTER(C.c_char_p)),
        ("n_materials", C.c_int32),
        ("mat_refractive_index", c_double_p),
        ("mat_comp_begin", c_int32_p),
        ("n_components", C.c_int32),
        ("comp_kind", c_int32_p),
        ("comp_coefficient", c_double_p),
        ("comp_abs_table", c_int32_p),
        ("comp_ems_table", c_int32_p),
        ("comp_quantum_yield", c_double_p),
        ("n_tables", C.c_int32),
        ("table_begin", c_int32_p),
        ("table_x", c_double_p),
        ("table_y", c_double_p),
    ]


class BatchDesc(C.Structure):
    _fields_ = [
        ("light_kind", C.c_int32),
        ("spectrum", C.c_int32),
        ("batch_id", C.c_uint32),
        ("reserved", C.c_uint32),
        ("wavelength_nm", C.c_double),
        ("mask_half", C.c_double * 2),
        ("mask_pose", C.c_double * 12),
        ("direction", C.c_double * 3),
        ("back_off", C.c_double),
        ("tilt_deg", C.c_double),
    ]


# the same layout as a NumPy structured dtype: the drivers fill thousands of descriptors column-wise (lights.py)
BATCH_DTYPE = np.dtype([("light_kind", "<i4"), ("spectrum", "<i4"), ("batch_id", "<u4"), ("reserved", "<u4"),
                        ("wavelength_nm", "<f8"), ("mask_half", "<f8", (2,)), ("mask_pose", "<f8", (12,)),
                        ("direction", "<f8", (3,)), ("back_off", "<f8"), ("tilt_deg", "<f8")], align=True)
assert BATCH_DTYPE.itemsize == C.sizeof(BatchDesc), (BATCH_DTYPE.itemsize, C.sizeof(BatchDesc))


def batch_pointer(table: np.ndarray):
    """``LscpmBatchDesc*`` view of a C-contiguous ``BATCH_DTYPE`` array (the array must outlive the call)."""
    assert table.dtype == BATCH_DTYPE and table.flags.c_contiguous
    return table.ctypes.data_as(C.POINTER(BatchDesc))


class SpectraDesc(C.Structure):
    _fields_ = [
        ("n_spectra", C.c_int32),
        ("begin", c_int32_p),
        ("x", c_double_p),
        ("y", c_double_p),
    ]


def _ptr(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class PackedScene:
    """Owns the numpy buffers behind a ``SceneDesc`` so the pointers stay valid."""

    def __init__(self, arrays):
        self.arrays = arrays
        keep = self._keep = {}
        for name, dtype in (("node_geom", np.int32), ("node_parent", np.int32), ("node_material", np.int32),
                            ("mat_comp_begin", np.int32), ("comp_kind", np.int32), ("comp_abs_table", np.int32),
                            ("comp_ems_table", np.int32), ("table_begin", np.int32)):
            keep[name] = np.ascontiguousarray(getattr(arrays, name), dtype=dtype)
        for name in ("node_size", "node_pose", "mat_refractive_index", "comp_coefficient", "comp_quantum_yield",
                     "table_x", "table_y"):
            keep[name] = np.ascontiguousarray(getattr(arrays, name), dtype=np.float64).reshape(-1)
        self._names = [str(n).encode("utf8") for n in arrays.node_name]
        self._name_array = (C.c_char_p * len(self._names))(*self._names)
        d = self.desc = SceneDesc()
        d.n_nodes = len(self._names)
        d.node_geom = _ptr(keep["node_geom"], C.c_int32)
        d.node_parent = _ptr(keep["node_parent"], C.c_int32)
        d.node_material = _ptr(keep["node_material"], C.c_int32)
        d.node_size = _ptr(keep["node_size"], C.c_double)
        d.node_pose = _ptr(keep["node_pose"], C.c_double)
        d.node_name = C.cast(self._name_array, C.POINTER(C.c_char_p))
        d.n_materials = len(keep["mat_refractive_index"])
        d.mat_refractive_index = _ptr(keep["mat_refractive_index"], C.c_double)
        d.mat_comp_begin = _ptr(keep["mat_comp_begin"], C.c_int32)
        d.n_components = len(keep["comp_kind"])
        d.comp_kind = _ptr(keep["comp_kind"], C.c_int32)
        d.comp_coefficient = _ptr(keep["comp_coefficient"], C.c_double)
        d.comp_abs_table = _ptr(keep["comp_abs_table"], C.c_int32)
        d.comp_ems_table = _ptr(keep["comp_ems_table"], C.c_int32)
        d.comp_quantum_yield = _ptr(keep["comp_quantum_yield"], C.c_double)
        d.n_tables = len(keep["table_begin"]) - 1
        d.table_begin = _ptr(keep["table_begin"], C.c_int32)
        d.table_x = _ptr(keep["table_x"], C.c_double)
        d.table_y = _ptr(keep["table_y"], C.c_double)


class PackedSpectra:
    """Spectra as ``LscpmSpectraDesc``: a sequence of (N_i, 2) knot arrays, or one (M, N, 2) array of M spectra with N
    knots each (the drivers' SPECTRL2 grid) which is packed without a Python loop."""

    def __init__(self, spectra):
        if isinstance(spectra, np.ndarray) and spectra.ndim == 3:
            m, k = spectra.shape[0], spectra.shape[1]
            self.n = m
            begin = (np.arange(m + 1, dtype=np.int64) * k).astype(np.int32)
            flat = np.ascontiguousarray(spectra, dtype=np.float64).reshape(-1, 2)
        else:
            self.n = len(spectra)
            begin = np.cumsum([0] + [len(s) for s in spectra]).astype(np.int32)
            flat = np.concatenate([np.asarray(s, dtype=np.float64).reshape(-1, 2) for s in spectra], axis=0) \
                if len(spectra) else np.zeros((0, 2))
        self._begin = np.ascontiguousarray(begin)
        self._x = np.ascontiguousarray(flat[:, 0])
        self._y = np.ascontiguousarray(flat[:, 1])
        d = self.desc = SpectraDesc()
        d.n_spectra = self.n
        d.begin = _ptr(self._begin, C.c_int32)
        d.x = _ptr(self._x, C.c_double)
        d.y = _ptr(self._y, C.c_double)


def fill_batch_row(row, light, batch_id: int = 0, spectrum_index: int = -1) -> None:
    """``flatten.LightArrays`` -> one row of a ``BATCH_DTYPE`` array."""
    row["light_kind"] = int(light.kind)
    row["spectrum"] = int(spectrum_index)
    row["batch_id"] = int(batch_id) & 0xFFFFFFFF
    row["reserved"] = 0
    row["wavelength_nm"] = float(light.wavelength_nm)
    row["mask_half"] = np.asarray(light.mask_half, dtype=np.float64).reshape(2)
    row["mask_pose"] = np.asarray(light.mask_pose, dtype=np.float64).reshape(12)
    row["direction"] = np.asarray(light.direction, dtype=np.float64).reshape(3)
    row["back_off"] = float(light.back_off)
    row["tilt_deg"] = float(light.tilt_deg)


def make_batch(light, batch_id: int = 0, spectrum_index: int = -1) -> BatchDesc:
    """``flatten.LightArrays`` -> ``LscpmBatchDesc``."""
    b = BatchDesc()
    b.light_kind = int(light.kind)
    b.spectrum = int(spectrum_index)
    b.batch_id = int(batch_id) & 0xFFFFFFFF
    b.reserved = 0
    b.wavelength_nm = float(light.wavelength_nm)
    b.mask_half[0], b.mask_half[1] = float(light.mask_half[0]), float(light.mask_half[1])
    pose = np.asarray(light.mask_pose, dtype=np.float64).reshape(12)
    for i in range(12):
        b.mask_pose[i] = float(pose[i])
    for i in range(3):
        b.direction[i] = float(light.direction[i])
    b.back_off = float(light.back_off)
    b.tilt_deg = float(light.tilt_deg)
    return b


_SIGNATURES = {
    "lscpm_abi_version": (C.c_int, []),
    "lscpm_last_error": (C.c_char_p, []),
    "lscpm_device_count": (C.c_int, []),
    "lscpm_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "lscpm_scene_create": (C.c_int, [C.POINTER(SceneDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "lscpm_scene_destroy": (C.c_int, [C.c_void_p]),
    "lscpm_scene_n_bins": (C.c_int, [C.c_void_p]),
    "lscpm_scene_bin_name": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
    "lscpm_scene_kernel_name": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "lscpm_desc_n_bins": (C.c_int, [C.POINTER(SceneDesc)]),
    "lscpm_desc_bin_name": (C.c_int, [C.POINTER(SceneDesc), C.c_int, C.c_char_p, C.c_int]),
    "lscpm_trace": (C.c_int, [C.c_void_p, C.POINTER(BatchDesc), C.c_int32, C.POINTER(SpectraDesc), C.c_uint64,
                              C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "lscpm_trace_host": (C.c_int, [C.c_void_p, C.POINTER(BatchDesc), C.c_int32, C.POINTER(SpectraDesc), C.c_uint64,
                                   C.c_uint64, C.c_uint64, c_int64_p]),
    "lscpm_plan_create": (C.c_int, [C.c_void_p, C.POINTER(BatchDesc), C.c_int32, C.POINTER(SpectraDesc), C.POINTER(C.c_void_p)]),
    "lscpm_plan_destroy": (C.c_int, [C.c_void_p]),
    "lscpm_trace_plan": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "lscpm_plan_h2d_bytes": (C.c_uint64, [C.c_void_p]),
    "lscpm_launch_count": (C.c_uint64, []),
    "lscpm_probe": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p,
                              c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "lscpm_emit": (C.c_int, [C.c_void_p, C.POINTER(BatchDesc), C.POINTER(SpectraDesc), C.c_uint64, C.c_int64,
                             C.c_uint64, c_double_p, c_double_p, c_double_p]),
    "lscpm_follow": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_uint32, C.c_uint64,
                               C.c_uint64, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_double_p, c_double_p,
                               c_double_p, c_int32_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Load ``liblscpm.so`` (once) and attach the prototypes; raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    target = path or os.environ.get("LSCPM_LIBRARY", LIB_PATH)
    if not os.path.exists(target):
        raise LscpmLibraryError(
            f"{target} not found: the CUDA library is not built.  Run `python -c \"import __graft_entry__ as g; "
            f"g.build()\"` (needs nvcc); there is no CPU fallback for the traced path.")
    lib = C.CDLL(target)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the library does not export what the header declares
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.lscpm_abi_version() != 1:
        raise LscpmLibraryError(f"ABI mismatch: library reports {lib.lscpm_abi_version()}, binding expects 1")
    if path is None:
        _lib = lib
    return lib


def check(lib: C.CDLL, code: int) -> None:
    if code != 0:
        message = lib.lscpm_last_error()
        raise LscpmError(code, message.decode("utf8", "replace") if message else "")
