"""GPU parity, check #1: deterministic per-ray interactions (intersection distance, interface, Fresnel
reflectivity, reflected and refracted directions) against the oracle for identical input rays.

Tolerance on the distance, asserted on EVERY ray (maximum, not a percentile):

    |t_gpu - t_oracle|  <=  1e-5 * t  +  ulp32(0.25 m) / |cos i|

The first term is the north star's 1e-5 relative.  The second is one unit of FP32 resolution of a plate coordinate
(2.98e-8 m at 0.25 m), taken along the surface normal: the kernels hold positions in FP32, so a hit point cannot be
placed better than that, and a displacement delta along the normal moves the distance by delta / |cos i|.  It only
matters for rays that start within micrometres of the surface they hit: scripts/probe_mismatch.py
(profiles/r2_probe_tail.txt, 3.1e6 rays, 7 scene / light cases) finds 0.15 % of the rays beyond 1e-5 relative, all of
them with t ~ 2e-5 m and |dt| ~ 7e-10 m (median), the IEEE-division build (-DLSCPM_PRECISE) giving the same numbers -
the tail is FP32 position rounding on sub-millimetre segments, not the SFU approximations - and a largest normalised
error of 0.53 against this bound."""
import numpy as np
import pytest

from helpers import standard_scene

pytestmark = pytest.mark.gpu

REL = 1e-5


def _harvest_rays(scene, n_photons, seed):
    """Rays at every step of oracle-followed photons: origins in air, plate, tube walls and mixture."""
    import pvtrace_oracle as po

    result = po.run(scene, n_photons, seed=seed, keep_histories=True)
    pos, direc = [], []
    for history in result.histories:
        for step in history[:-1]:
            pos.append(step.position)
            direc.append(step.direction)
    return np.array(pos), np.array(direc)


ULP_PLATE = float(np.spacing(np.float32(0.25)))


def _harvest_rays_c(arrays, native, n_photons, seed):
    """Rays at every step of photons followed by the C oracle (large populations: ~9 rays per photon)."""
    import c_oracle
    from lscpm_b200 import backend

    oracle = c_oracle.COracle(native, native.PackedScene(arrays))
    batches, _ = backend.pack_batches([arrays.light])
    pos, direc, wl = oracle.emit(batches[0], n_photons, seed=seed)
    hist = oracle.follow(pos, direc, wl, seed=seed + 1, max_steps=48)
    P, D = [], []
    for i in range(len(pos)):
        k = min(hist["n_steps"][i], 48)
        keep = np.isin(hist["event"][i, :k - 1], (0, 1, 2, 4))
        P.append(hist["pos"][i, :k - 1][keep])
        D.append(hist["dir"][i, :k - 1][keep])
    return np.concatenate(P), np.concatenate(D)


def _check_rays(arrays, native, pos, direc, min_rays):
    import c_oracle
    from lscpm_b200 import backend

    # identical FP32-representable inputs on both sides
    pos = pos.astype(np.float32).astype(np.float64)
    direc = direc.astype(np.float32).astype(np.float64)
    direc /= np.linalg.norm(direc, axis=1, keepdims=True)
    ref = c_oracle.COracle(native, native.PackedScene(arrays)).probe(pos, direc)
    gpu = backend.probe_rays(arrays, pos, direc, device=0)

    # rays whose origin sits within float rounding of a surface are ambiguous after the cast; keep clear ones
    clear = ref["distance"] > 1e-6
    same_iface = (gpu["hit"] == ref["hit"]) & (gpu["container"] == ref["container"]) & (gpu["adjacent"] == ref["adjacent"])
    mismatches = int(np.sum(clear & ~same_iface))
    assert mismatches <= 3, f"interface bookkeeping differs on {mismatches} of {int(clear.sum())} clear rays"
    ok = clear & same_iface & (ref["hit"] > 0)
    assert ok.sum() > min_rays
    t = ref["distance"][ok]
    dt = np.abs(gpu["distance"][ok] - t)
    cosi = np.abs(np.einsum("ij,ij->i", ref["normal"][ok], direc[ok]))
    bound = REL * t + ULP_PLATE / np.maximum(cosi, 1e-12)
    worst = int(np.argmax(dt / bound))
    assert np.all(dt <= bound), (f"distance outside 1e-5*t + ulp/|cos i| on {int(np.sum(dt > bound))} rays; worst: t {t[worst]:.3e} "
                                 f"dt {dt[worst]:.3e} cos i {cosi[worst]:.3e} normalised {dt[worst] / bound[worst]:.2f}")
    d_rel = dt / t
    assert np.quantile(d_rel, 0.99) < REL and np.median(d_rel) < 2e-6
    r_err = np.abs(gpu["reflectivity"][ok] - ref["reflectivity"][ok])
    steep = np.abs(ref["reflectivity"][ok] - 1.0) > 1e-9
    assert np.quantile(r_err[steep], 0.99) < 1e-4
    tir_ref = ref["reflectivity"][ok] >= 1.0
    tir_gpu = gpu["reflectivity"][ok] >= 1.0
    assert np.mean(tir_ref == tir_gpu) > 0.999
    both = ~tir_ref & ~tir_gpu
    for key in ("reflected", "refracted", "normal"):
        err = np.linalg.norm(gpu[key][ok][both] - ref[key][ok][both], axis=1)
        assert np.quantile(err, 0.99) < 5e-5, (key, np.quantile(err, 0.99))


@pytest.mark.parametrize("tilt,elevation", [(0, 90), (40, 50)])
def test_probe_matches_oracle(tilt, elevation, native, built_library):
    from lscpm_b200 import flatten

    scene = standard_scene(tilt=tilt, elevation=elevation, wavelength=585.0)
    arrays = flatten.flatten_scene(scene)
    pos, direc = _harvest_rays(scene, 400, seed=11)
    _check_rays(arrays, native, pos, direc, 1000)


@pytest.mark.parametrize("label,kwargs", [
    ("oblique", dict(tilt=0, elevation=25, azimuth=100, wavelength=585.0)),
    ("diffuse", dict(tilt=20, wavelength=600.0, diffuse=True)),
    ("pv_both", dict(tilt=0, elevation=70, azimuth=120, wavelength=585.0, add_bottom_PV=True, add_side_PV=True)),
    ("dense46", dict(tilt=0, elevation=60, azimuth=140, wavelength=585.0, capillary_count=46, capillary_pitch=0.01)),
])
def test_probe_matches_oracle_on_every_ray_of_a_large_population(label, kwargs, native, built_library):
    """~1.5e5 rays per case; the maximum error is asserted (the p99 gate could not see 1 ray in 500)."""
    from lscpm_b200 import flatten

    arrays = flatten.flatten_scene(standard_scene(**kwargs))
    pos, direc = _harvest_rays_c(arrays, native, 20_000, seed=5)
    _check_rays(arrays, native, pos, direc, 50_000)


def test_known_answer_vectors(native, built_library):
    """Analytic vectors of SURVEY.md Appendix E through the CUDA probe (tilt 0: world z = local z - 0.004)."""
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, wavelength=585.0))
    top = 0.0  # plate top face in world coordinates
    th = np.radians([40.0, 60.0, 80.0])
    pos = np.array([[0.05, 0.02, top + 0.01]] * 3)
    direc = np.stack([np.sin(th), np.zeros(3), -np.cos(th)], axis=1)
    pos = pos - direc * 0.0 + np.array([0, 0, 0])  # start 1 cm above, hit the top face
    out = backend.probe_rays(arrays, pos, direc, device=0)
    np.testing.assert_allclose(out["reflectivity"], [0.0430463296, 0.0858444380, 0.384207216], rtol=2e-5)
    np.testing.assert_allclose(out["refracted"][:, 0], [0.434315952, 0.585152300, 0.665410644], rtol=2e-5)
    np.testing.assert_allclose(out["refracted"][:, 2], [-0.900760597, -0.810923416, -0.746477511], rtol=2e-5)
    # E6: vertical ray through capillary 4 from just under the top face
    p = np.array([[-0.1055, 0.0, top - 1e-9]])
    d = np.array([[0.0, 0.0, -1.0]])
    out = backend.probe_rays(arrays, p, d, device=0)
    assert arrays.node_name[out["hit"][0]] == "Capillary_PFA_4"
    np.testing.assert_allclose(out["distance"][0], 0.00249329623, rtol=1e-5)
