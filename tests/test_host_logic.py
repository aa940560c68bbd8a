"""Host-side logic that needs no GPU: the flattener, the light recognition, the tally naming contract, the C-ABI
library's exports, and the loud failure of the traced path without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import am15_like_photon_spectrum, standard_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flatten_standard_scene():
    from lscpm_b200 import flatten, tally

    arrays = flatten.flatten_scene(standard_scene(tilt=40, elevation=50, wavelength=555.0))
    assert arrays.n_nodes == 34
    assert arrays.node_name[0] == "World (air)" and arrays.node_name[1].startswith("LSC-PM")
    assert arrays.node_name[2] == "Capillary_PFA_0" and arrays.node_name[3] == "Reaction_mixture_0"   # pre-order
    assert list(arrays.node_parent[:4]) == [-1, 0, 1, 2]
    assert arrays.node_geom[0] == tally.GEOM_SPHERE and arrays.node_geom[1] == tally.GEOM_BOX
    assert np.allclose(arrays.node_size[1], [0.47, 0.47, 0.008])
    assert np.allclose(arrays.node_size[2], [0.47, 0.0015875, 0])
    assert np.allclose(arrays.node_size[3], [0.47, 0.00079375, 0])
    # materials: world, plate, shared PFA, one per reaction cylinder; one shared Reactor component
    assert len(arrays.mat_refractive_index) == 19
    assert sorted(set(arrays.mat_refractive_index)) == [1.0, 1.34, 1.344, 1.48]
    assert list(arrays.comp_kind).count(tally.COMP_REACTOR) == 16   # CSR rows repeat the shared component
    assert list(arrays.comp_kind).count(tally.COMP_LUMINOPHORE) == 1
    # tables de-duplicated by content, header row dropped (reference scene_creator.py:75-80)
    assert len(arrays.table_begin) - 1 == 3
    assert sorted(np.diff(arrays.table_begin)) == [450, 451, 1300]
    # plate pose: tilt about +y with the top face through the origin
    pose = arrays.node_pose[1].reshape(4, 4)
    t = np.radians(40)
    assert np.allclose(pose[:3, 2], [np.sin(t), 0, np.cos(t)])
    assert np.allclose(pose[:3, 3], [-np.sin(t) * 0.004, 0, -np.cos(t) * 0.004])
    # capillary axis (local z) maps to the plate's y
    assert np.allclose(np.abs(arrays.node_pose[2].reshape(4, 4)[:3, 2]), [0, 1, 0])


def test_flatten_light_direct_diffuse_spectrum():
    from lscpm_b200 import flatten

    light = flatten.flatten_scene(standard_scene(tilt=40, elevation=50, azimuth=200, wavelength=555.0)).light
    assert light.kind == flatten.LIGHT_DIRECT and light.back_off == pytest.approx(1.0)
    el, az = np.radians(50), np.radians(180 - 200)
    sun = np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    assert np.allclose(light.direction, -sun)
    assert np.allclose(light.mask_pose[:, 3], 0)                       # anchors stay on the plate face
    t = np.radians(40)
    assert np.allclose(light.mask_pose[:, 0], [np.cos(t), 0, -np.sin(t)])
    assert light.wavelength_nm == 555.0 and light.spectrum is None and light.wavelength_callable is None

    light = flatten.flatten_scene(standard_scene(tilt=25, diffuse=True, spectrum=am15_like_photon_spectrum())).light
    assert light.kind == flatten.LIGHT_DIFFUSE and light.tilt_deg == 25
    assert light.spectrum.shape == (27, 2) and light.spectrum[0, 0] == 360 and light.spectrum[-1, 0] == 690

    import random

    from lscpm_b200.miniplant.scene_creator import create_direct_scene

    rng = random.Random(1)
    light = flatten.flatten_scene(create_direct_scene(solar_spectrum_function=lambda: rng.choice([500, 600]))).light
    assert light.wavelength_callable is not None                       # unknown callable: host emission path


def test_unsupported_light_is_reported():
    from lscpm_b200 import flatten
    from lscpm_b200.compat.pvtrace_api import Light, Node
    from lscpm_b200.miniplant.scene_creator import _create_scene_common

    scene = _create_scene_common(0, Node(name="custom", light=Light()))
    with pytest.raises(flatten.UnsupportedSceneError):
        flatten.flatten_light(scene)


def test_bin_names_python_and_c_agree(native, built_library):
    from lscpm_b200 import flatten

    for kw in (dict(), dict(include_dye=False)):
        arrays = flatten.flatten_scene(standard_scene(**kw))
        packed = native.PackedScene(arrays)
        n = built_library.lscpm_desc_n_bins(C.byref(packed.desc))
        names = arrays.bin_names()
        assert n == len(names) == 3 + 33 * 6 + (2 if kw == {} else 1) + 32
        buf = C.create_string_buffer(256)
        for b in range(n):
            assert built_library.lscpm_desc_bin_name(C.byref(packed.desc), b, buf, 256) > 0
            assert buf.value.decode() == names[b]
        assert names[:3] == ["kill", "exit/miss", "exit/entry_reflect"]
        assert "react/Reaction_mixture_7/0" in names and "exit/LSC-PM Reactor 47x47 cm^2/+z" in names


def test_oracle_bin_layout_agrees(native):
    import c_oracle
    from lscpm_b200 import flatten

    arrays = flatten.flatten_scene(standard_scene())
    assert c_oracle.COracle(native, native.PackedScene(arrays)).n_bins == len(arrays.bin_names())


def test_library_exports_every_declared_symbol(native, built_library):
    """Every function include/lscpm.h declares is exported by the shared library (no compute calls here)."""
    header = open(os.path.join(ROOT, "include", "lscpm.h")).read()
    declared = set(re.findall(r"\b(lscpm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(built_library, name), f"{name} declared in include/lscpm.h but not exported"
    assert declared == set(native.EXPORTED_SYMBOLS), declared ^ set(native.EXPORTED_SYMBOLS)
    assert built_library.lscpm_abi_version() == 1


def test_library_is_built_for_sm_100a():
    import subprocess

    lib = os.path.join(ROOT, "lsc-pm_solar_miniplant_b200", "csrc", "liblscpm.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_invalid_descriptions_are_rejected(native, built_library):
    from lscpm_b200 import flatten

    arrays = flatten.flatten_scene(standard_scene())
    arrays.node_material = arrays.node_material.copy()
    arrays.node_material[5] = 99
    packed = native.PackedScene(arrays)
    assert built_library.lscpm_desc_n_bins(C.byref(packed.desc)) == -1          # LSCPM_EINVAL
    assert b"material" in built_library.lscpm_last_error()
    handle = C.c_void_p()
    assert built_library.lscpm_scene_create(C.byref(packed.desc), 0, C.byref(handle)) == -1


def test_traced_path_fails_loudly_without_a_gpu(native, built_library):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lscpm_b200 import backend, flatten
    from miniplant.simulation_runner import run_direct_simulation

    assert built_library.lscpm_device_count() == 0
    arrays = flatten.flatten_scene(standard_scene())
    packed = native.PackedScene(arrays)
    handle = C.c_void_p()
    rc = built_library.lscpm_scene_create(C.byref(packed.desc), 0, C.byref(handle))
    assert rc == -3 and b"no CUDA device" in built_library.lscpm_last_error()   # LSCPM_ECUDA, no CPU fallback
    with pytest.raises(backend.NoCudaDeviceError):
        run_direct_simulation(num_photons=10)


def test_missing_library_is_an_error(native):
    with pytest.raises(native.LscpmLibraryError):
        native.load("/nonexistent/liblscpm.so")


def test_render_with_workers_raises_before_anything_else():
    from miniplant.simulation_runner import _common_simulation_runner

    with pytest.raises(RuntimeError, match="renderer"):
        _common_simulation_runner(object(), 10, render=True, workers=2)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or execute it."""
    pkg = os.path.join(ROOT, "lsc-pm_solar_miniplant_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f)).read()
                for needle in ("pvtrace_oracle", "c_oracle", "liblscpm_oracle", "oracle_trace", "lscpm_oracle"):
                    assert needle not in text, f"{f} refers to {needle}"


def test_spectral_unit_conversion():
    from lscpm_b200.compat.pvtrace_api import Distribution
    from miniplant.utils import irradiance_to_photon_flux, photon_energy, spectral_distribution_to_photon_distribution

    assert photon_energy(555) == pytest.approx(3.5792e-19, rel=1e-4)
    assert irradiance_to_photon_flux(1.0, 500) == pytest.approx(500e-9 / (6.62607015e-34 * 299792458.0) / 6.02214076e23)
    d = spectral_distribution_to_photon_distribution(Distribution(np.array([400.0, 500.0]), np.array([1.0, 1.0])), 1800)
    assert d._y[1] / d._y[0] == pytest.approx(500 / 400)
