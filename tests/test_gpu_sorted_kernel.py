"""The block-sorted kernel (K2s) and the pool kernel (K2p, two configurations) against the lane-per-photon kernel (GPU).

A photon's fate depends only on (seed, batch id, photon index) and all kernels run the same device functions on the same
Philox draws, so their tallies must be IDENTICAL, bin for bin and counter for counter - whatever the scheduling.  The
library picks one by scene and launch size (lscpm.cu: lscpm_scene_create / lscpm_trace_plan); LSCPM_TRACE_KERNEL forces one.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import am15_like_photon_spectrum, standard_scene  # noqa: E402

CASES = {
    "config1_585nm": dict(tilt=0, elevation=90, wavelength=585.0),
    "config2_spectrum": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()),
    "tilted_oblique": dict(tilt=40, elevation=35, azimuth=140, spectrum=am15_like_photon_spectrum()),
    "diffuse": dict(tilt=30, diffuse=True, spectrum=am15_like_photon_spectrum()),
    "no_dye": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum(), include_dye=False),
    "pv_both": dict(tilt=0, elevation=60, spectrum=am15_like_photon_spectrum(), add_bottom_PV=True, add_side_PV=True),
    "pv_backlit": dict(tilt=40, elevation=20, azimuth=0, wavelength=555.0, add_bottom_PV=True, add_side_PV=True),
    "dense46": dict(tilt=0, elevation=60, azimuth=140, wavelength=585.0, capillary_count=46, capillary_pitch=0.01),
}


def _tallies(arrays, which, n_batches, photons, seed, photon_begin=0):
    from lscpm_b200 import backend
    old = os.environ.get("LSCPM_TRACE_KERNEL")
    os.environ["LSCPM_TRACE_KERNEL"] = which
    try:
        scene = backend.SceneHandle(arrays, 0)   # a fresh handle: the choice is made at scene creation
    finally:
        if old is None:
            del os.environ["LSCPM_TRACE_KERNEL"]
        else:
            os.environ["LSCPM_TRACE_KERNEL"] = old
    plan = backend.Plan(scene, [arrays.light] * n_batches, list(range(10, 10 + n_batches)))
    tallies = plan.new_tallies()
    plan.trace_into(tallies, photons, seed, photon_begin=photon_begin)
    import torch
    torch.cuda.synchronize()
    return tallies.cpu().numpy(), scene.n_bins


@pytest.mark.parametrize("label", sorted(CASES))
def test_sorted_kernel_is_bit_identical(label):
    from lscpm_b200 import flatten
    arrays = flatten.flatten_scene(standard_scene(**CASES[label]))
    for n_batches, photons in ((1, 1), (3, 257), (2, 70_001), (5, 300_000)):
        a, n_bins = _tallies(arrays, "lanes", n_batches, photons, seed=1234)
        b, _ = _tallies(arrays, "sorted", n_batches, photons, seed=1234)
        assert (b[:, :n_bins].sum(axis=1) == photons).all(), f"{label}: sorted kernel lost or duplicated photons"
        np.testing.assert_array_equal(a, b, err_msg=f"{label} batches={n_batches} photons={photons}")


POOL_CASES = sorted(CASES)   # scenes with outside boxes included: their pool kernel has an air phase and carries `aux` in the record


@pytest.mark.parametrize("which", ["pool120", "pool96", "pool80", "pool64"])
@pytest.mark.parametrize("label", POOL_CASES)
def test_pool_kernel_is_bit_identical(label, which):
    from lscpm_b200 import flatten
    arrays = flatten.flatten_scene(standard_scene(**CASES[label]))
    for n_batches, photons in ((1, 1), (3, 257), (2, 70_001), (5, 300_000)):
        a, n_bins = _tallies(arrays, "lanes", n_batches, photons, seed=4321)
        b, _ = _tallies(arrays, which, n_batches, photons, seed=4321)
        assert (b[:, :n_bins].sum(axis=1) == photons).all(), f"{label}: pool kernel lost or duplicated photons"
        np.testing.assert_array_equal(a, b, err_msg=f"{label} {which} batches={n_batches} photons={photons}")


def test_pool_kernel_photon_offset_and_many_small_batches():
    from lscpm_b200 import flatten
    arrays = flatten.flatten_scene(standard_scene(tilt=20, elevation=50, spectrum=am15_like_photon_spectrum()))
    a, n_bins = _tallies(arrays, "lanes", 400, 37, seed=5, photon_begin=(1 << 33) + 12345)
    for which in ("pool120", "pool96", "pool80", "pool64"):
        b, _ = _tallies(arrays, which, 400, 37, seed=5, photon_begin=(1 << 33) + 12345)
        assert (b[:, :n_bins].sum(axis=1) == 37).all()
        np.testing.assert_array_equal(a, b)


def test_default_kernel_of_the_standard_scene_is_the_pool_kernel_and_agrees_with_lanes():
    """What a user gets without any environment variable, against the lane kernel."""
    from lscpm_b200 import backend, flatten
    arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()))
    a, n_bins = _tallies(arrays, "lanes", 4, 250_000, seed=99)
    assert "LSCPM_TRACE_KERNEL" not in os.environ
    scene = backend.SceneHandle(arrays, 0)
    assert scene.kernel_name.startswith("trace_pool_kernel<ext=0, flight=1, warps=24, records=96>"), scene.kernel_name
    plan = backend.Plan(scene, [arrays.light] * 4, list(range(10, 14)))
    tallies = plan.new_tallies()
    plan.trace_into(tallies, 250_000, 99)
    np.testing.assert_array_equal(a, tallies.cpu().numpy())
    # a plate without luminophore stays on the lane kernel; scenes with outside boxes use the pool kernel too (plate flights included: no box shares a z face)
    no_dye = backend.SceneHandle(flatten.flatten_scene(standard_scene(include_dye=False)), 0)
    assert no_dye.kernel_name.startswith("trace_kernel<ext=0, flight=0"), no_dye.kernel_name
    pv = backend.SceneHandle(flatten.flatten_scene(standard_scene(add_bottom_PV=True)), 0)
    assert pv.kernel_name.startswith("trace_pool_kernel<ext=1, flight=1"), pv.kernel_name


def test_sorted_kernel_photon_offset_and_many_small_batches():
    """Work items shorter than a block (many batches of a few photons) and a photon-index offset beyond 2^32."""
    from lscpm_b200 import flatten
    arrays = flatten.flatten_scene(standard_scene(tilt=20, elevation=50, spectrum=am15_like_photon_spectrum()))
    a, n_bins = _tallies(arrays, "lanes", 400, 37, seed=5, photon_begin=(1 << 33) + 12345)
    b, _ = _tallies(arrays, "sorted", 400, 37, seed=5, photon_begin=(1 << 33) + 12345)
    assert (b[:, :n_bins].sum(axis=1) == 37).all()
    np.testing.assert_array_equal(a, b)
