"""Device plumbing between the scene-level API and the C-ABI (``include/lscpm.h``).

torch is used for what it is good at here — device memory, streams, ``torch.distributed`` — and for
nothing else: every photon is traced by the hand-written kernels in ``csrc/``.  There is no CPU
path: without a CUDA device (or without the built library) these functions raise.

Replaces the body of ``_common_simulation_runner`` (reference ``miniplant/simulation_runner.py:30-66``):
``scene.emit`` + ``photon_tracer.follow`` per photon + ``Counter(finals)``.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import weakref
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _native as nat
from . import flatten as _flatten
from . import tally as _tally
from .compat import pvtrace_api as _api


class NoCudaDeviceError(RuntimeError):
    pass


def _torch():
    import torch

    return torch


def device_index(device=None) -> int:
    torch = _torch()
    if not torch.cuda.is_available():
        raise NoCudaDeviceError(
            "no CUDA device is visible: the LSC-PM photon tracer runs only on the GPU (sm_100a); "
            "the CPU restatement in oracle/ is test infrastructure and is not used as a fallback")
    if device is None:
        return torch.cuda.current_device()
    return torch.device(device).index if not isinstance(device, int) else device


# --------------------------------------------------------------------------------------
# scene handles (cached by content: the drivers rebuild an identical scene for every time point,
# reference miniplant/simulation_runner.py:97-105)
# --------------------------------------------------------------------------------------


class SceneHandle:
    def __init__(self, arrays: _flatten.SceneArrays, device: int):
        self.lib = nat.load()
        self.arrays = arrays
        self.packed = nat.PackedScene(arrays)
        self.device = device
        handle = C.c_void_p()
        nat.check(self.lib, self.lib.lscpm_scene_create(C.byref(self.packed.desc), device, C.byref(handle)))
        self.handle = handle
        self.n_bins = self.lib.lscpm_scene_n_bins(handle)
        self.bin_names = [self._bin_name(b) for b in range(self.n_bins)]
        self.stride = self.n_bins + nat.N_COUNTERS
        # released when the handle is collected; at interpreter exit the process teardown reclaims device memory,
        # streams and events anyway, and CUDA calls made while other libraries unwind are not worth the risk
        weakref.finalize(self, self.lib.lscpm_scene_destroy, handle).atexit = False

    @property
    def kernel_name(self) -> str:
        """The trace kernel launches on this scene use (``lscpm_scene_kernel_name``)."""
        buf = C.create_string_buffer(256)
        nat.check(self.lib, min(self.lib.lscpm_scene_kernel_name(self.handle, buf, 256), 0))
        return buf.value.decode("utf8")

    def _bin_name(self, b: int) -> str:
        buf = C.create_string_buffer(256)
        n = self.lib.lscpm_scene_bin_name(self.handle, b, buf, 256)
        if n < 0:
            nat.check(self.lib, n)
        return buf.value.decode("utf8")


_scene_cache: dict = {}


def _scene_key(arrays: _flatten.SceneArrays, device: int) -> str:
    h = hashlib.sha1()
    h.update(repr((device, arrays.node_name)).encode())
    for name in ("node_geom", "node_parent", "node_material", "node_size", "node_pose", "mat_refractive_index",
                 "mat_comp_begin", "comp_kind", "comp_coefficient", "comp_abs_table", "comp_ems_table",
                 "comp_quantum_yield", "table_begin", "table_x", "table_y"):
        h.update(np.ascontiguousarray(getattr(arrays, name)).tobytes())
    return h.hexdigest()


def scene_handle(arrays: _flatten.SceneArrays, device=None) -> SceneHandle:
    dev = device_index(device)
    pinned = arrays.__dict__.setdefault("_handles", {})   # the same SceneArrays object asked again: no hashing
    handle = pinned.get(dev)
    if handle is not None:
        return handle
    key = _scene_key(arrays, dev)
    handle = _scene_cache.get(key)
    if handle is None:
        if len(_scene_cache) > 64:
            _scene_cache.clear()
        handle = _scene_cache[key] = SceneHandle(arrays, dev)
    pinned[dev] = handle
    return handle


# --------------------------------------------------------------------------------------
# batches
# --------------------------------------------------------------------------------------


def batch_table(lights, batch_ids: Optional[Sequence[int]] = None):
    """Anything that describes batches -> ``lights.BatchTable``: a table passes through (``batch_ids`` override its
    ids), a sequence of ``LightArrays`` is packed row by row with its spectra de-duplicated."""
    from .lights import BatchTable

    if isinstance(lights, BatchTable):
        if batch_ids is None:
            return lights
        desc = lights.desc.copy()
        desc["batch_id"] = np.asarray(batch_ids, dtype=np.uint32)
        return BatchTable(desc, lights.spectra)
    n = len(lights)
    desc = np.zeros(n, dtype=nat.BATCH_DTYPE)
    spectra, index = [], {}
    for i, light in enumerate(lights):
        if light.wavelength_callable is not None:
            raise _flatten.UnsupportedSceneError("host-sampled wavelength callables go through trace_rays()")
        spec = -1
        if light.spectrum is not None:
            knots = np.ascontiguousarray(light.spectrum, dtype=np.float64)
            key = (id(light.spectrum), knots.shape) if n > 64 else knots.tobytes()
            if key not in index:
                index[key] = len(spectra)
                spectra.append(knots)
            spec = index[key]
        nat.fill_batch_row(desc[i], light, batch_id=(batch_ids[i] if batch_ids is not None else i), spectrum_index=spec)
    return BatchTable(desc, spectra or None)


def pack_batches(lights, batch_ids: Optional[Sequence[int]] = None):
    """-> (``LscpmBatchDesc*`` over the table's rows, ``PackedSpectra``, the table that owns the memory)."""
    table = batch_table(lights, batch_ids)
    spectra = nat.PackedSpectra(table.spectra if table.spectra is not None else [])
    return _BatchPointer(table), spectra


class _BatchPointer:
    """ctypes view of a table's rows that keeps the table alive; indexable like the ctypes array it replaces."""

    def __init__(self, table):
        self.table = table
        self._as_parameter_ = nat.batch_pointer(table.desc)

    def __len__(self):
        return len(self.table)

    def __getitem__(self, i):
        return self._as_parameter_[i]


class Plan:
    """Batch descriptors and spectra resident on the device (``lscpm_plan_create``)."""

    def __init__(self, scene: SceneHandle, lights, batch_ids=None):
        self.scene = scene
        self._batches, self._spectra = pack_batches(lights, batch_ids)
        self.n_batches = len(self._batches)
        handle = C.c_void_p()
        nat.check(scene.lib, scene.lib.lscpm_plan_create(scene.handle, self._batches, self.n_batches,
                                                         C.byref(self._spectra.desc), C.byref(handle)))
        self.handle = handle
        weakref.finalize(self, scene.lib.lscpm_plan_destroy, handle).atexit = False

    def new_tallies(self):
        torch = _torch()
        return torch.zeros((self.n_batches, self.scene.stride), dtype=torch.int64, device=f"cuda:{self.scene.device}")

    def trace_into(self, tallies, photons_per_batch: int, seed: int, photon_begin: int = 0) -> None:
        """Asynchronous launch on torch's current stream; ADDS into the int64 device tensor ``tallies``."""
        torch = _torch()
        assert tallies.dtype == torch.int64 and tallies.is_cuda and tallies.is_contiguous()
        assert tuple(tallies.shape) == (self.n_batches, self.scene.stride)
        stream = torch.cuda.current_stream(tallies.device).cuda_stream
        lib = self.scene.lib
        nat.check(lib, lib.lscpm_trace_plan(self.handle, int(photon_begin), int(photons_per_batch),
                                            int(seed) & 0xFFFFFFFFFFFFFFFF, C.c_void_p(tallies.data_ptr()),
                                            C.c_void_p(stream)))


def fresh_seed() -> int:
    """The reference is unseeded (global numpy RNG); an unspecified seed draws fresh entropy."""
    return int.from_bytes(os.urandom(8), "little")


def trace_lights(arrays: _flatten.SceneArrays, lights, photons_per_batch: int,
                 seed: Optional[int] = None, device=None, batch_ids=None, photon_begin: int = 0) -> np.ndarray:
    """Trace ``photons_per_batch`` photons for every light configuration (a sequence of ``LightArrays`` or a
    ``lights.BatchTable``); returns int64 ``[n_batches, n_bins + N_COUNTERS]`` on the host.

    One call of ``lscpm_trace_host``: host descriptors in, host tallies out, through the scene's pinned staging buffer
    and device scratch — no allocation, no torch tensor, one stream synchronisation."""
    scene = scene_handle(arrays, device)
    seed = fresh_seed() if seed is None else int(seed)
    batches, spectra = pack_batches(lights, batch_ids)
    out = np.zeros((len(batches), scene.stride), dtype=np.int64)
    if len(batches) and int(photons_per_batch) > 0:
        nat.check(scene.lib, scene.lib.lscpm_trace_host(scene.handle, batches, len(batches), C.byref(spectra.desc),
                                                        int(photon_begin), int(photons_per_batch),
                                                        seed & 0xFFFFFFFFFFFFFFFF, out.ctypes.data_as(nat.c_int64_p)))
    return out


def trace_rays(arrays: _flatten.SceneArrays, positions, directions, wavelengths, seed: Optional[int] = None,
               device=None, batch_id: int = 0, photon_begin: int = 0, max_steps: int = 0) -> dict:
    """Trace caller-supplied rays (world frame) on the GPU: ``lscpm_follow``.  With ``max_steps > 0`` the
    per-step history is returned as well."""
    scene = scene_handle(arrays, device)
    lib = scene.lib
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    direc = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    wl = np.ascontiguousarray(wavelengths, dtype=np.float64).reshape(-1)
    n, m = len(pos), int(max_steps)
    seed = fresh_seed() if seed is None else int(seed)
    out = dict(n_steps=np.zeros(n, np.int32), final_bin=np.zeros(n, np.int32))
    dp, ip = nat.c_double_p, nat.c_int32_p
    if m > 0:
        out.update(event=np.full((n, m), -1, np.int32), node=np.full((n, m), -1, np.int32), pos=np.zeros((n, m, 3)),
                   dir=np.zeros((n, m, 3)), wavelength=np.zeros((n, m)))
        extra = [out["event"].ctypes.data_as(ip), out["node"].ctypes.data_as(ip), out["pos"].ctypes.data_as(dp),
                 out["dir"].ctypes.data_as(dp), out["wavelength"].ctypes.data_as(dp)]
    else:
        extra = [None, None, None, None, None]
    nat.check(lib, lib.lscpm_follow(scene.handle, n, pos.ctypes.data_as(dp), direc.ctypes.data_as(dp),
                                    wl.ctypes.data_as(dp), batch_id, photon_begin, seed & 0xFFFFFFFFFFFFFFFF, m,
                                    out["n_steps"].ctypes.data_as(ip), *extra, out["final_bin"].ctypes.data_as(ip)))
    out["bin_names"] = scene.bin_names
    return out


def probe_rays(arrays: _flatten.SceneArrays, positions, directions, device=None) -> dict:
    """Deterministic next-interface query for world-frame rays: ``lscpm_probe``."""
    scene = scene_handle(arrays, device)
    lib = scene.lib
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    direc = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    n = len(pos)
    out = dict(hit=np.zeros(n, np.int32), container=np.zeros(n, np.int32), adjacent=np.zeros(n, np.int32),
               face=np.zeros(n, np.int32), distance=np.zeros(n), point=np.zeros((n, 3)), normal=np.zeros((n, 3)),
               reflectivity=np.zeros(n), reflected=np.zeros((n, 3)), refracted=np.zeros((n, 3)))
    dp, ip = nat.c_double_p, nat.c_int32_p
    nat.check(lib, lib.lscpm_probe(scene.handle, n, pos.ctypes.data_as(dp), direc.ctypes.data_as(dp),
                                   out["hit"].ctypes.data_as(ip), out["container"].ctypes.data_as(ip),
                                   out["adjacent"].ctypes.data_as(ip), out["face"].ctypes.data_as(ip),
                                   out["distance"].ctypes.data_as(dp), out["point"].ctypes.data_as(dp),
                                   out["normal"].ctypes.data_as(dp), out["reflectivity"].ctypes.data_as(dp),
                                   out["reflected"].ctypes.data_as(dp), out["refracted"].ctypes.data_as(dp)))
    return out


def emit_rays(arrays: _flatten.SceneArrays, light: _flatten.LightArrays, n: int, seed: int = 0, device=None,
              batch_id: int = 0, photon_begin: int = 0):
    """The first ``n`` rays the device generates for a light (world frame): ``lscpm_emit``."""
    scene = scene_handle(arrays, device)
    lib = scene.lib
    batches, spectra = pack_batches([light], [batch_id])
    pos, direc, wl = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
    dp = nat.c_double_p
    nat.check(lib, lib.lscpm_emit(scene.handle, batches, C.byref(spectra.desc), photon_begin, n,
                                  int(seed) & 0xFFFFFFFFFFFFFFFF, pos.ctypes.data_as(dp), direc.ctypes.data_as(dp),
                                  wl.ctypes.data_as(dp)))
    return pos, direc, wl


# --------------------------------------------------------------------------------------
# scene-level entry points used by the miniplant mirror and the pvtrace-compatible shims
# --------------------------------------------------------------------------------------


def simulate_scene(scene, num_photons: int, seed: Optional[int] = None, device=None) -> _tally.TallyResult:
    """Trace ``num_photons`` photons of the scene's light to termination and tally them.

    Recognised lights (the reference's direct and diffuse sources with constant or tabulated spectra) are
    generated on the device.  Any other light — arbitrary user callables — is emitted on the host through the
    scene's own ``emit`` (exactly the rays the reference would trace) and traced on the device."""
    arrays = _flatten.flatten_scene(scene, with_light=False)
    num_photons = int(num_photons)
    try:
        light = _flatten.flatten_light(scene)
        host_emit = light.wavelength_callable is not None
    except _flatten.UnsupportedSceneError:
        light, host_emit = None, True
    if num_photons <= 0:
        handle = scene_handle(arrays, device)
        return _tally.TallyResult(handle.bin_names, np.zeros(handle.n_bins, np.int64), 0, 0)
    if not host_emit:
        row = trace_lights(arrays, [light], num_photons, seed=seed, device=device)[0]
        handle = scene_handle(arrays, device)
        return _tally.TallyResult(handle.bin_names, row[:handle.n_bins].copy(), int(row[handle.n_bins]),
                                  int(row[handle.n_bins + 1]))
    rays = list(scene.emit(num_photons))
    out = trace_rays(arrays, [r.position for r in rays], [r.direction for r in rays], [r.wavelength for r in rays],
                     seed=seed, device=device)
    handle = scene_handle(arrays, device)
    counts = np.bincount(out["final_bin"], minlength=handle.n_bins).astype(np.int64)
    return _tally.TallyResult(handle.bin_names, counts, 0, 0)


def capture_paths(scene, num_paths: int, seed: Optional[int] = None, device=None, max_steps: int = 256) -> list:
    """Photon paths for visual debugging (reference ``photon_path`` / ``renderer.add_ray_path``,
    ``miniplant/simulation_runner.py:34-43,55``): the first ``num_paths`` rays of the scene's light, followed on the
    GPU with per-step records.  Each path is a tuple of ``Ray`` objects, GENERATE ray first."""
    arrays = _flatten.flatten_scene(scene, with_light=False)
    rays = list(scene.emit(int(num_paths)))
    if not rays:
        return []
    out = trace_rays(arrays, [r.position for r in rays], [r.direction for r in rays], [r.wavelength for r in rays],
                     seed=seed, device=device, max_steps=max_steps)
    paths = []
    for i, ray in enumerate(rays):
        steps = [ray]
        for k in range(min(int(out["n_steps"][i]), max_steps)):
            steps.append(_api.Ray(position=tuple(out["pos"][i, k]), direction=tuple(out["dir"][i, k]),
                                  wavelength=float(out["wavelength"][i, k]), source=getattr(ray, "source", None)))
        paths.append(tuple(steps))
    return paths


def simulate_final_events(scene, num_rays: int, **kwargs) -> list:
    """``Scene.simulate`` shape: one history per photon whose last entry carries the terminal ``Event``
    (the reference reads only ``photon[-1][1]``, ``miniplant/simulation_runner.py:63``)."""
    result = simulate_scene(scene, num_rays, **kwargs)
    finals = []
    for name, count in zip(result.names, result.counts):
        if not count:
            continue
        if name.startswith("react/"):
            event = _api.Event.REACT
        elif name.startswith("absorb/"):
            event = _api.Event.ABSORB
        elif name == "kill":
            event = _api.Event.KILL
        else:
            event = _api.Event.EXIT
        finals.extend([[(None, event)]] * int(count))
    return finals


_EVENT_OF_CODE = {0: _api.Event.GENERATE, 1: _api.Event.REFLECT, 2: _api.Event.TRANSMIT, 3: _api.Event.ABSORB,
                  4: _api.Event.EMIT, 5: _api.Event.EXIT, 6: _api.Event.KILL, 7: _api.Event.REACT}


def follow_ray(scene, ray, maxsteps: int = 1000, seed: Optional[int] = None, device=None, **_ignored) -> list:
    """``photon_tracer.follow(scene, ray)`` on the GPU: the ``[(Ray, Event), ...]`` history of one ray."""
    arrays = _flatten.flatten_scene(scene, with_light=False)
    cap = min(int(maxsteps), 1000) + 2
    out = trace_rays(arrays, [ray.position], [ray.direction], [ray.wavelength], seed=seed, device=device, max_steps=cap)
    history = [(ray, _api.Event.GENERATE)]
    n = int(out["n_steps"][0])
    last = ray
    for k in range(min(n, cap)):
        last = _api.Ray(position=tuple(out["pos"][0, k]), direction=tuple(out["dir"][0, k]),
                        wavelength=float(out["wavelength"][0, k]), source=getattr(ray, "source", None))
        history.append((last, _EVENT_OF_CODE[int(out["event"][0, k])]))
    final = out["bin_names"][int(out["final_bin"][0])]
    if final.startswith("exit/") and history[-1][1] is not _api.Event.EXIT:
        history.append((last, _api.Event.EXIT))
    return history


def next_hit_ray(scene, ray):
    """``next_hit(scene, ray)`` on the GPU: ``(hit_node, (container, adjacent), point, distance)`` or None."""
    arrays = _flatten.flatten_scene(scene, with_light=False)
    out = probe_rays(arrays, [ray.position], [ray.direction])
    nodes = [n for n in _iter_geometry_nodes(scene.root)]
    hit = int(out["hit"][0])
    if hit < 0:
        return None
    container, adjacent = int(out["container"][0]), int(out["adjacent"][0])
    return (nodes[hit], (nodes[container], nodes[adjacent] if adjacent >= 0 else None), tuple(out["point"][0]),
            float(out["distance"][0]))


def _iter_geometry_nodes(node) -> Iterable:
    if getattr(node, "geometry", None) is not None:
        yield node
    for child in node.children:
        yield from _iter_geometry_nodes(child)
