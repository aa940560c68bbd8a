"""GPU parity, check #1: deterministic per-ray interactions (intersection distance, interface, Fresnel
reflectivity, reflected and refracted directions) against the oracle for identical input rays, 1e-5 relative."""
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


@pytest.mark.parametrize("tilt,elevation", [(0, 90), (40, 50)])
def test_probe_matches_oracle(tilt, elevation, native, built_library):
    import c_oracle
    from lscpm_b200 import backend, flatten

    scene = standard_scene(tilt=tilt, elevation=elevation, wavelength=585.0)
    arrays = flatten.flatten_scene(scene)
    pos, direc = _harvest_rays(scene, 400, seed=11)
    # identical FP32-representable inputs on both sides
    pos = pos.astype(np.float32).astype(np.float64)
    direc = direc.astype(np.float32).astype(np.float64)
    direc /= np.linalg.norm(direc, axis=1, keepdims=True)
    ref = c_oracle.COracle(native, native.PackedScene(arrays)).probe(pos, direc)
    gpu = backend.probe_rays(arrays, pos, direc, device=0)

    # rays whose origin sits within float rounding of a surface are ambiguous after the cast; keep clear ones
    clear = ref["distance"] > 1e-6
    same_iface = (gpu["hit"] == ref["hit"]) & (gpu["container"] == ref["container"]) & (gpu["adjacent"] == ref["adjacent"])
    frac = np.mean(same_iface[clear])
    assert frac > 0.9995, f"interface bookkeeping agrees on {frac:.4f} of rays"
    ok = clear & same_iface & (ref["hit"] > 0)
    assert ok.sum() > 1000
    d_rel = np.abs(gpu["distance"][ok] - ref["distance"][ok]) / ref["distance"][ok]
    tol = REL   # holds for the tilted plate too: the world->plate rotation is applied in double on the host
    assert np.quantile(d_rel, 0.99) < tol, f"distance rel err p99 {np.quantile(d_rel, 0.99):.2e}"
    assert np.median(d_rel) < 2e-6
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
