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


# ---- round 2: column-wise host stage ----------------------------------------------------------------------------------


def test_fast_csv_is_byte_identical_to_pandas(tmp_path):
    import pandas as pd
    from lscpm_b200.miniplant._csv import write_csv

    rng = np.random.default_rng(3)
    for tz in ("Europe/Amsterdam", "Australia/Brisbane", "UTC", None):
        index = pd.date_range("2020-03-28 22:00", periods=400, freq="1800s", tz=tz)      # spans a DST change
        frame = pd.DataFrame({"apparent_elevation": rng.normal(30, 20, 400), "azimuth": rng.uniform(0, 360, 400),
                              "simulation_direct": rng.integers(0, 121, 400) / 120.0, "count": rng.integers(0, 9, 400),
                              "direct_reacted": rng.uniform(0, 1e-3, 400)}, index=index)
        frame.iloc[5, 0] = np.nan
        frame.iloc[7, 2] = 0.0
        frame.iloc[9, 4] = 1e-320
        cols = ("apparent_elevation", "azimuth", "simulation_direct", "count", "direct_reacted")
        write_csv(frame, tmp_path / "fast.csv", cols)
        frame.to_csv(tmp_path / "pandas.csv", columns=cols)
        assert (tmp_path / "fast.csv").read_bytes() == (tmp_path / "pandas.csv").read_bytes(), tz
    # anything the fast writer does not format itself goes through pandas
    frame = pd.DataFrame({"a": ["x", "y"], "b": [1.0, 2.0]}, index=pd.RangeIndex(2))
    write_csv(frame, tmp_path / "f.csv", ("a", "b"))
    frame.to_csv(tmp_path / "p.csv", columns=("a", "b"))
    assert (tmp_path / "f.csv").read_bytes() == (tmp_path / "p.csv").read_bytes()


def test_batch_tables_equal_the_per_row_descriptors():
    """lights.direct_batches / diffuse_batches (column-wise) against backend.batch_table over per-row LightArrays."""
    from lscpm_b200 import _native as nat, backend, lights
    from lscpm_b200.workloads import AM15_LIKE_W, SPECTRL2_SLICE_NM

    assert nat.BATCH_DTYPE.itemsize == 176
    rng = np.random.default_rng(5)
    n = 37
    elevation, azimuth = rng.uniform(1, 89, n), rng.uniform(60, 300, n)
    flux = AM15_LIKE_W * SPECTRL2_SLICE_NM * rng.uniform(0.5, 1.5, (n, 1)) * rng.uniform(0.9, 1.1, (n, len(SPECTRL2_SLICE_NM)))
    flux[11] = flux[3]      # a repeated spectrum is stored once
    table = lights.BatchTable.concatenate([lights.direct_batches(-15.0, elevation, azimuth, flux=flux, x=SPECTRL2_SLICE_NM),
                                           lights.diffuse_batches(-15.0, flux=flux[:5], x=SPECTRL2_SLICE_NM)])
    rows = [lights.direct_light(-15.0, e, a, np.stack([SPECTRL2_SLICE_NM, f], axis=1)) for e, a, f in zip(elevation, azimuth, flux)]
    rows += [lights.diffuse_light(-15.0, np.stack([SPECTRL2_SLICE_NM, f], axis=1)) for f in flux[:5]]
    packed = backend.batch_table(rows)
    assert len(table) == len(packed) == n + 5 and len(table.spectra) == n - 1 + 5
    for field in ("light_kind", "batch_id", "wavelength_nm", "mask_half", "mask_pose", "direction", "back_off", "tilt_deg"):
        assert np.array_equal(table.desc[field], packed.desc[field]), field
    for i in range(len(table)):
        assert np.array_equal(table.spectra[table.desc["spectrum"][i]], packed.spectra[packed.desc["spectrum"][i]])
        light = table[i]
        assert light.kind == rows[i].kind and np.array_equal(light.spectrum, rows[i].spectrum)
        assert np.array_equal(light.direction, rows[i].direction) and np.array_equal(light.mask_pose, rows[i].mask_pose)
    sub = table.take([40, 2, 7])
    assert list(sub.desc["batch_id"]) == [40, 2, 7]
    spectra = nat.PackedSpectra(table.spectra)
    assert spectra.n == len(table.spectra) and spectra._begin[-1] == len(table.spectra) * len(SPECTRL2_SLICE_NM)


def test_year_batches_column_wise_and_row_wise_agree():
    import pandas as pd
    from lscpm_b200 import backend, lights
    from lscpm_b200.compat.pvtrace_api import Distribution
    from lscpm_b200.miniplant.full_simulation import year_batches
    from lscpm_b200.miniplant.solar_data import solar_data_from_positions

    data = solar_data_from_positions([30.0, 45.0, 60.0], [150.0, 180.0, 210.0], 30, index=pd.RangeIndex(3))
    fast = year_batches(30, data)
    assert isinstance(fast, lights.BatchTable) and len(fast) == 6
    data.at[1, "direct_spectrum"] = Distribution(np.array([360.0, 500.0, 690.0]), np.array([1.0, 2.0, 1.0]))   # other abscissae
    slow = year_batches(30, data)
    assert isinstance(slow, list) and len(slow) == 6
    a, b = fast.desc, backend.batch_table(slow).desc
    for field in ("light_kind", "batch_id", "mask_pose", "direction", "back_off", "tilt_deg"):
        assert np.array_equal(a[field], b[field]), field


def test_lazy_distribution_behaves_like_distribution():
    from lscpm_b200.compat.pvtrace_api import Distribution, LazyDistribution
    from lscpm_b200.workloads import am15_like_photon_spectrum

    spec = am15_like_photon_spectrum()
    a, b = Distribution(spec[:, 0], spec[:, 1]), LazyDistribution(spec[:, 0], spec[:, 1])
    u = np.linspace(0, 1, 101)
    assert np.array_equal(a.sample(u), b.sample(u)) and np.array_equal(a.lookup(spec[:, 0]), b.lookup(spec[:, 0]))
    assert np.array_equal(a(500.0), b(500.0)) and isinstance(b, Distribution) and b.hist is False


def test_committed_ncu_summary_belongs_to_these_kernel_sources():
    """bench.py replays the counters of profiles/r2_final_metrics.json in its roofline object only when the summary was
    captured from the sources the library is built from.  This pins that the committed pair is
    consistent: whoever edits a kernel re-captures the profile (scripts/run_final_profile_r2.sh, make_final_profile.py)."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_digest", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    metrics, reason = bench.committed_profile()
    assert metrics is not None, reason
    assert metrics["source_sha256"] == bench.source_digest()
    assert metrics["kernel"].startswith("void trace_pool_kernel<0, 1, 24, 120>")
    assert metrics["photons_per_launch"] == 64_000_000
