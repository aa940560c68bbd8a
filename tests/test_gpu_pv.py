"""Row f3: the photovoltaic variants (reference miniplant/scene_creator.py:99-206, simulation_runner.py:45-53,68-79):
boxes outside the plate make the surrounding air a medium with geometry.  Tallies and per-ray interfaces against the
generic oracle, and the reference's polymorphic 5-tuple return."""
import numpy as np
import pytest

from helpers import assert_counts_agree, assert_tallies_agree, grouped, standard_scene

pytestmark = pytest.mark.gpu

CASES = {
    "bottom": dict(tilt=40, elevation=50, wavelength=555.0, add_bottom_PV=True),
    "side": dict(tilt=40, elevation=50, wavelength=555.0, add_side_PV=True),
    "both_585": dict(tilt=0, elevation=90, wavelength=585.0, add_bottom_PV=True, add_side_PV=True),
    "both_backlit": dict(tilt=40, elevation=20, azimuth=0, wavelength=555.0, add_bottom_PV=True, add_side_PV=True),
    "both_no_dye": dict(tilt=10, elevation=60, wavelength=610.0, include_dye=False, add_bottom_PV=True, add_side_PV=True),
}


def _oracle(arrays, native):
    import c_oracle

    return c_oracle.COracle(native, native.PackedScene(arrays))


@pytest.mark.parametrize("case", sorted(CASES))
def test_pv_tallies_match_oracle(case, native, built_library):
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(standard_scene(**CASES[case]))
    n = 300_000
    row = backend.trace_lights(arrays, [arrays.light], n, seed=3, device=0)[0]
    batches, _ = backend.pack_batches([arrays.light])
    ref = _oracle(arrays, native).trace(batches[0], photon_count=n, seed=4)
    names = arrays.bin_names()
    nb = len(names)
    assert row[:nb].sum() == n and ref[:nb].sum() == n
    assert_tallies_agree(names, row[:nb], ref[:nb], label=case, resample=lambda: backend.trace_lights(
        arrays, [arrays.light], n, seed=5, device=0)[0][:nb])
    assert abs(row[nb] - ref[nb]) / ref[nb] < 0.01           # interaction events (the absorbing box counts two)
    if "backlit" not in case:
        assert sum(c for name, c in zip(names, row) if "PV" in name and name.startswith("absorb/")) > 0.05 * n


def test_box_on_a_z_face_rules_plate_flights_out(native, built_library):
    """Plate flights fold the reflections at the z faces with the test against the world's medium; the library enables
    them for scenes with outside boxes only when no box shares one of those faces.  The reference's bottom cells hang
    centimetres below the plate (flights on); pushed up against the plate they are optically coupled to it (n = 3.4 beyond the
    face: no total internal reflection there) and the cell march must run - tallies against the oracle either way."""
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, wavelength=585.0, add_bottom_PV=True))
    assert "flight=1" in backend.SceneHandle(arrays, 0).kernel_name
    i = arrays.node_name.index("bottomPV")
    plate = int(arrays.node_parent[i])
    # poses are world poses; the plate is level (tilt 0), so only z matters
    touching = arrays.node_pose[plate, 11] - (arrays.node_size[plate, 2] + arrays.node_size[i, 2]) / 2
    assert arrays.node_pose[i, 11] < touching - 0.01            # as built: an air gap of centimetres
    arrays.node_pose[i, 11] = touching                          # top face of the cells = bottom face of the plate
    scene = backend.SceneHandle(arrays, 0)
    assert scene.kernel_name.startswith("trace_pool_kernel<ext=1, flight=0"), scene.kernel_name
    n = 300_000
    row = backend.trace_lights(arrays, [arrays.light], n, seed=3, device=0)[0]
    batches, _ = backend.pack_batches([arrays.light])
    ref = _oracle(arrays, native).trace(batches[0], photon_count=n, seed=4)
    names = arrays.bin_names()
    nb = len(names)
    assert row[:nb].sum() == n and ref[:nb].sum() == n
    assert_tallies_agree(names, row[:nb], ref[:nb], label="coupled cells", resample=lambda: backend.trace_lights(
        arrays, [arrays.light], n, seed=5, device=0)[0][:nb])
    assert abs(row[nb] - ref[nb]) / ref[nb] < 0.01
    # the coupled cells drain the trapped light: several times the share the hanging cells see
    coupled = sum(c for name, c in zip(names, row) if name.startswith("absorb/bottomPV"))
    assert coupled > 0.3 * n, coupled / n


def test_pv_probe_matches_oracle(native, built_library):
    import pvtrace_oracle as po
    from lscpm_b200 import backend, flatten

    scene = standard_scene(tilt=0, elevation=90, wavelength=585.0, add_bottom_PV=True, add_side_PV=True)
    arrays = flatten.flatten_scene(scene)
    result = po.run(scene, 300, seed=2, keep_histories=True)
    pos, direc = [], []
    for history in result.histories:
        for step in history[:-1]:
            pos.append(step.position)
            direc.append(step.direction)
    pos = np.array(pos).astype(np.float32).astype(np.float64)
    direc = np.array(direc).astype(np.float32).astype(np.float64)
    direc /= np.linalg.norm(direc, axis=1, keepdims=True)
    ref = _oracle(arrays, native).probe(pos, direc)
    gpu = backend.probe_rays(arrays, pos, direc, device=0)
    clear = ref["distance"] > 1e-5
    same = (gpu["hit"] == ref["hit"]) & (gpu["container"] == ref["container"]) & (gpu["adjacent"] == ref["adjacent"])
    assert np.mean(same[clear]) > 0.999
    in_air_to_pv = clear & same & np.isin(ref["hit"], [i for i, name in enumerate(arrays.node_name) if "PV" in name])
    assert in_air_to_pv.sum() > 20
    lending = clear & same & (ref["hit"] == 1) & np.isin(ref["adjacent"], [i for i, name in enumerate(arrays.node_name) if "PV" in name])
    assert lending.sum() == 0      # a photon leaving the plate meets the world's medium, whatever box lies beyond the face
    leaving = clear & same & (ref["hit"] == 1) & (ref["container"] == 1)
    assert leaving.sum() > 100 and np.all(ref["adjacent"][leaving] == 0)
    ok = clear & same & (ref["hit"] > 0)
    rel = np.abs(gpu["distance"][ok] - ref["distance"][ok]) / ref["distance"][ok]
    assert np.quantile(rel, 0.99) < 1e-5
    r_err = np.abs(gpu["reflectivity"][ok] - ref["reflectivity"][ok])
    assert np.quantile(r_err[ref["reflectivity"][ok] < 1.0], 0.99) < 1e-4


def test_runner_returns_the_reference_tuple(built_library):
    """reference simulation_runner.py:127-139: res[-2] = bottom PV count, res[-1] = side PV count."""
    from miniplant.simulation_runner import run_direct_simulation

    res = run_direct_simulation(tilt_angle=40, solar_elevation=50, solar_azimuth=180, add_bottom_PV=True, add_side_PV=True,
                                num_photons=20_000, seed=1)
    assert isinstance(res, tuple) and len(res) == 5
    fraction, scene, photon_path, bottom, side = res
    # C oracle at 4e5 photons: reacted 0.295, bottom PV 0.166, side PVs 0.120 (total internal reflection survives over the
    # bottom PV; at the edge faces the side PVs share, the box's face sorts first half of the time and lends its index)
    assert 0.27 < fraction < 0.32 and 0.15 * 20_000 < bottom < 0.185 * 20_000 and 0.105 * 20_000 < side < 0.135 * 20_000
    assert hasattr(scene, "light_nodes") and photon_path == []
    plain = run_direct_simulation(tilt_angle=40, solar_elevation=50, solar_azimuth=180, num_photons=2000, seed=1)
    assert isinstance(plain, float)


def test_photon_path_capture(built_library):
    """Row f4: photon paths for the reference's photon_path slot / renderer (simulation_runner.py:34-43,55)."""
    from lscpm_b200 import backend
    from lscpm_b200.compat.pvtrace_api import Ray
    from miniplant.scene_creator import create_direct_scene
    from miniplant.simulation_runner import run_direct_simulation

    res = run_direct_simulation(tilt_angle=0, solar_elevation=90, add_bottom_PV=True, num_photons=5000, seed=2, capture=25)
    paths = res[2]
    assert len(paths) == 25 and all(isinstance(step, Ray) for path in paths for step in path)
    assert all(len(path) >= 2 for path in paths)
    first = paths[0]
    assert abs(first[0].position[2] - 1.0) < 1e-9 and abs(first[1].position[2]) < 1e-6     # 1 m above, then on the top face
    scene = create_direct_scene(tilt_angle=0, solar_elevation=90)
    plain = backend.capture_paths(scene, 10, seed=3)
    assert len(plain) == 10
