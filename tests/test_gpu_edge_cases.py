"""Edge cases of the hot path on the GPU: empty and ragged inputs, many tiny batches, 64-bit photon indices, scene
variants of the LSC-PM family (no capillaries, no inner cylinder, dense channel grid) against the generic oracle, and
scenes outside the family (rejected loudly, never traced as something else)."""
import numpy as np
import pytest

from helpers import am15_like_photon_spectrum, assert_counts_agree, assert_tallies_agree, grouped, standard_scene

pytestmark = pytest.mark.gpu


def _oracle_row(arrays, native, n, seed=5):
    import c_oracle
    from lscpm_b200 import backend

    oracle = c_oracle.COracle(native, native.PackedScene(arrays))
    batches, _ = backend.pack_batches([arrays.light])
    spectra = [arrays.light.spectrum] if arrays.light.spectrum is not None else None
    return oracle.trace(batches[0], spectra=spectra, photon_count=n, seed=seed)


def test_zero_and_tiny_photon_counts(built_library):
    from lscpm_b200 import backend, flatten
    from miniplant.simulation_runner import run_direct_simulation

    arrays = flatten.flatten_scene(standard_scene())
    for n in (1, 2, 31, 33, 120):
        row = backend.trace_lights(arrays, [arrays.light], n, seed=n, device=0)[0]
        assert row[:-2].sum() == n
    assert run_direct_simulation(num_photons=0) == 0.0
    assert 0.0 <= run_direct_simulation(num_photons=1, seed=3) <= 1.0


def test_many_tiny_batches_like_the_yearly_driver(native, built_library):
    """16 k time points x 120 photons (reference full_simulation.py:53): every row sums to 120; pooled tallies follow
    the oracle."""
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(standard_scene(tilt=40, elevation=50, wavelength=555.0))
    nb, ppb = 4000, 120
    rows = backend.trace_lights(arrays, [arrays.light] * nb, ppb, seed=8, device=0)
    names = arrays.bin_names()
    assert rows.shape == (nb, len(names) + 2)
    assert np.all(rows[:, :len(names)].sum(axis=1) == ppb)
    assert len({tuple(r) for r in rows[:50]}) > 40            # distinct batch ids -> distinct random streams
    pooled = rows.sum(axis=0)
    ref = _oracle_row(arrays, native, nb * ppb)
    g, r = grouped(names, pooled[:len(names)]), grouped(names, ref[:len(names)])
    keys = sorted(set(g) | set(r))
    assert_counts_agree(keys, [g.get(k, 0) for k in keys], [r.get(k, 0) for k in keys], label="tiny batches pooled")


def test_photon_indices_beyond_32_bits(built_library):
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(standard_scene())
    base = (1 << 33) + 12345
    whole = backend.trace_lights(arrays, [arrays.light], 6000, seed=4, device=0, photon_begin=base)
    parts = sum(backend.trace_lights(arrays, [arrays.light], 2000, seed=4, device=0, photon_begin=base + 2000 * i) for i in range(3))
    assert np.array_equal(whole, parts)
    low = backend.trace_lights(arrays, [arrays.light], 6000, seed=4, device=0, photon_begin=12345)
    assert not np.array_equal(whole, low)                      # the high counter word matters


def _custom_scene(n_channels, pitch, with_inner=True, r_outer=0.0015875, r_inner=0.00079375, tilt=0.0, wavelength=585.0,
                  plate=(0.47, 0.47, 0.008), cells=False):
    """A member of the LSC-PM family built with the same object API the reference uses."""
    from lscpm_b200.compat import pvtrace_api as pv
    from lscpm_b200.miniplant import scene_creator as sc
    from lscpm_b200.miniplant.utils import LightPosition, VectorInverter
    from lscpm_b200.workloads import ConstantWavelength

    spectra = sc._spectra()
    world = pv.Node(name="World (air)", geometry=pv.Sphere(radius=10.0, material=pv.Material(refractive_index=1.0)))
    sun = pv.spherical_to_cart(np.deg2rad(20.0), np.deg2rad(15.0))
    light = pv.Node(name="Solar Light", light=pv.Light(wavelength=ConstantWavelength(wavelength), direction=VectorInverter(sun),
                                                       position=LightPosition(tilt_angle=tilt)), parent=world)
    light.translate(sun)
    dye = pv.Luminophore(coefficient=np.array(spectra["lr305_abs"]), emission=np.array(spectra["lr305_ems"]),
                         quantum_yield=0.95, phase_function=pv.isotropic)
    reactor = pv.Node(name="plate", geometry=pv.Box(size=plate, material=pv.Material(
        refractive_index=1.48, components=[pv.Absorber(coefficient=0.1), dye])), parent=world)
    if cells:   # photovoltaic absorbers as the reference places them: one under the plate behind an air gap, one strip on an edge
        opaque = pv.Material(refractive_index=3.4, components=[pv.Absorber(coefficient=1e10)])
        pv.Node(name="bottomPV", geometry=pv.Box(size=plate, material=opaque), parent=reactor).translate((0, 0, -0.025))
        pv.Node(name="sidePV1", geometry=pv.Box(size=(plate[0], 0.01, plate[2]), material=opaque),
                parent=reactor).translate((0, 0.5 * plate[1] + 0.005, 0))
    mix = pv.Reactor(np.array(spectra["mb_abs"]))
    tube_geom = pv.Cylinder(length=plate[1], radius=r_outer, material=pv.Material(refractive_index=1.34, components=[pv.Absorber(coefficient=0.1)]))
    x0 = -0.5 * pitch * (n_channels - 1)
    for k in range(n_channels):
        tube = pv.Node(name=f"tube_{k}", geometry=tube_geom, parent=reactor)
        if with_inner:
            pv.Node(name=f"mix_{k}", geometry=pv.Cylinder(length=plate[1], radius=r_inner, material=pv.Material(
                refractive_index=1.344, components=[mix])), parent=tube)
        tube.rotate(np.radians(90), (1, 0, 0))
        tube.translate((x0 + pitch * k, 0, 0))
    reactor.rotate(np.radians(tilt), (0, 1, 0))
    reactor.translate((-np.sin(np.deg2rad(tilt)) * 0.5 * plate[2], 0, -np.cos(np.deg2rad(tilt)) * 0.5 * plate[2]))
    return pv.Scene(world)


@pytest.mark.parametrize("label,kw", [
    ("bare_plate", dict(n_channels=0, pitch=0.03)),
    ("tubes_without_mixture", dict(n_channels=16, pitch=0.03, with_inner=False)),
    ("single_capillary", dict(n_channels=1, pitch=0.03)),
    ("dense_grid_46_channels", dict(n_channels=46, pitch=0.01, wavelength=555.0)),
    ("fat_tubes_tilted", dict(n_channels=8, pitch=0.05, r_outer=0.003, r_inner=0.002, tilt=25.0)),
    # outside boxes with plate flights (pool kernel, fifth list) on family members other than the reference's plate
    ("bare_plate_with_cells", dict(n_channels=0, pitch=0.03, tilt=15.0, cells=True)),
    ("dense_grid_with_cells", dict(n_channels=46, pitch=0.01, wavelength=555.0, cells=True)),
])
def test_scene_family_variants_match_oracle(label, kw, native, built_library):
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(_custom_scene(**kw))
    if kw.get("cells"):
        assert backend.scene_handle(arrays, 0).kernel_name.startswith("trace_pool_kernel<ext=1, flight=1"), backend.scene_handle(arrays, 0).kernel_name
    n = 200_000
    row = backend.trace_lights(arrays, [arrays.light], n, seed=21, device=0)[0]
    ref = _oracle_row(arrays, native, n)
    names = arrays.bin_names()
    assert row[:len(names)].sum() == n
    assert_tallies_agree(names, row[:len(names)], ref[:len(names)], label=label,
                         resample=lambda: backend.trace_lights(arrays, [arrays.light], n, seed=22, device=0)[0][:len(names)])
    assert abs(row[len(names)] - ref[len(names)]) / ref[len(names)] < 0.02


def test_scenes_outside_the_family_are_rejected(native, built_library):
    from lscpm_b200 import backend, flatten
    from miniplant.simulation_runner import run_direct_simulation

    from lscpm_b200.compat import pvtrace_api as pv

    # a box that overlaps the plate is not an "outside box"
    scene = _custom_scene(n_channels=2, pitch=0.03)
    plate = scene.root.children[1]
    pv.Node(name="intruder", geometry=pv.Box(size=(0.1, 0.1, 0.004), material=pv.Material(
        refractive_index=3.4, components=[pv.Absorber(coefficient=1e10)])), parent=plate)
    with pytest.raises(native.LscpmError) as err:
        backend.trace_lights(flatten.flatten_scene(scene), [flatten.flatten_light(scene)], 10, seed=1, device=0)
    assert err.value.code == -2 and "overlaps" in str(err.value)
    # capillaries shorter than the plate: caps inside the plate are a different topology
    scene = _custom_scene(n_channels=2, pitch=0.03)
    for node in scene.root.children[1].children:
        node.geometry = pv.Cylinder(length=0.4, radius=0.0015875, material=node.geometry.material)
    with pytest.raises(native.LscpmError):
        backend.trace_lights(flatten.flatten_scene(scene), [flatten.flatten_light(scene)], 10, seed=1, device=0)


@pytest.mark.parametrize("dye,side_pv", [(False, False), (True, False), (True, True)])
def test_coincident_end_caps_match_oracle(dye, side_pv, native, built_library):
    """Coincident faces: the capillary end caps lie in the plate's +-y faces (and the side PV boxes on its edges).
    pvtrace orders such tied intersections by float64 round-off; oracle and kernels draw the order at random (DESIGN.md
    section 4).  A strip light over the last 2 mm before the +y edge, shining towards it, sends ~0.5 % of its photons
    out through the caps instead of ~0.03 %: every bin - among them exit/<tube or mixture>/<cap>, 64 of them - and every
    family must agree with the generic oracle."""
    import dataclasses

    import c_oracle
    from lscpm_b200 import backend, flatten

    kw = dict(add_side_PV=True) if side_pv else {}
    arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, wavelength=610.0, include_dye=dye, **kw))
    pose = np.array(arrays.light.mask_pose, dtype=float).copy()
    pose[1, 3] = 0.234                                                   # strip y in [0.233, 0.235] on the top face
    direction = np.array([0.1, np.sin(np.radians(55.0)), -np.cos(np.radians(55.0))])
    light = dataclasses.replace(arrays.light, mask_half=np.array([0.235, 0.001]), mask_pose=pose,
                                direction=direction / np.linalg.norm(direction))
    n = 1_000_000
    row = backend.trace_lights(arrays, [light], n, seed=21, device=0)[0]
    oracle = c_oracle.COracle(native, native.PackedScene(arrays))
    batches, _ = backend.pack_batches([light])
    ref = oracle.trace(batches[0], photon_count=n, seed=22)
    names = arrays.bin_names()
    nb = len(names)
    assert row[:nb].sum() == n and ref[:nb].sum() == n
    caps = np.array([name.endswith("cap") for name in names])
    if not side_pv:   # (with a side PV on the edge the photons that leave through a cap end in the box)
        assert ref[:nb][caps].sum() > 0.003 * n                           # the caps really are exercised
    assert_tallies_agree(names, row[:nb], ref[:nb], label=f"caps dye={dye} pv={side_pv}",
                         resample=lambda: backend.trace_lights(arrays, [light], n, seed=23, device=0)[0][:nb])
    assert abs(row[nb] - ref[nb]) / ref[nb] < 0.005                       # interaction events


def test_scenes_with_many_bins_share_the_kernel(built_library):
    """Two live scene handles whose per-warp tally rows both need the shared-memory opt-in (> 48 KB per block) but
    different amounts: the opt-in limit belongs to the kernel, not to a handle, so creating the smaller one must not
    break launches of the larger one."""
    from lscpm_b200 import backend, flatten

    big = flatten.flatten_scene(standard_scene(tilt=0, elevation=80, wavelength=585.0, capillary_count=46, capillary_pitch=0.01))
    small = flatten.flatten_scene(standard_scene(tilt=0, elevation=80, wavelength=585.0, capillary_count=40, capillary_pitch=0.0115))
    n = 50_000
    first = backend.trace_lights(big, [big.light], n, seed=8, device=0)[0]
    other = backend.trace_lights(small, [small.light], n, seed=8, device=0)[0]
    again = backend.trace_lights(big, [big.light], n, seed=8, device=0)[0]
    assert first[:-2].sum() == n and other[:-2].sum() == n and np.array_equal(first, again)


def test_plate_array_with_dense_channel_grid_one_plate_per_batch(native, built_library):
    """BASELINE configs[4] in small: a 3 x 3 array of dense-grid plates (46 capillaries at 10 mm).  The plates of an array
    are optically independent (photons are independent; the reference has no multi-plate scene), so every
    plate is one batch with its own tilt, sun position, RNG stream and tally row; plates that share a tilt share a launch.
    1e6 photons over the array, every plate's bins and pooled families against the oracle tracing the same plate."""
    import c_oracle
    from lscpm_b200 import backend, flatten, lights
    from lscpm_b200.workloads import AM15_LIKE_W, SPECTRL2_SLICE_NM

    photons = 111_112                                  # x 9 plates = 1e6
    flux = (AM15_LIKE_W * SPECTRL2_SLICE_NM)[None, :].repeat(3, axis=0)
    plate = 0
    for row, tilt in enumerate((0.0, 20.0, 40.0)):     # one row of the array per tilt
        scene = standard_scene(tilt=tilt, elevation=60, azimuth=180, spectrum=am15_like_photon_spectrum(), capillary_count=46, capillary_pitch=0.01)
        arrays = flatten.flatten_scene(scene)
        names = arrays.bin_names()
        elevation = np.array([25.0, 50.0, 75.0]) + row
        azimuth = np.array([120.0, 180.0, 235.0]) - 3 * row
        table = lights.direct_batches(tilt, elevation, azimuth, flux=flux, x=SPECTRL2_SLICE_NM, first_batch_id=3 * row)
        rows = backend.trace_lights(arrays, table, photons, seed=31, device=0)
        assert rows.shape == (3, len(names) + 2) and np.all(rows[:, :len(names)].sum(axis=1) == photons)
        oracle = c_oracle.COracle(native, native.PackedScene(arrays))
        for i in range(3):
            light = table[i]
            batches, _ = backend.pack_batches([light], [plate])
            ref = oracle.trace(batches[0], spectra=[light.spectrum], photon_count=photons, seed=900 + plate)
            assert_tallies_agree(names, rows[i][:len(names)], ref[:len(names)], label=f"plate {plate} (tilt {tilt})",
                                 resample=lambda i=i: backend.trace_lights(arrays, table, photons, seed=77, device=0)[i][:len(names)])
            plate += 1
    assert plate == 9


def test_bare_plate_folded_flights_at_high_statistics(native, built_library):
    """The plate without capillaries has no coincident faces and no adjacency question: for it the generic oracle is a
    plain box billiard with Fresnel faces, while the GPU folds every run of total internal reflections into one plate
    flight.  2e7 photons (most of the re-emitted light is trapped and walks to the edges): every bin within 4.5 sigma, no
    re-draw, and the events per photon - which count every folded reflection - to 1.5e-3 (3.5 sigma)."""
    from helpers import bins_outside_band
    from lscpm_b200 import backend, flatten

    arrays = flatten.flatten_scene(_custom_scene(n_channels=0, pitch=0.03, tilt=10.0))
    handle = backend.scene_handle(arrays, 0)
    assert "flight=1" in handle.kernel_name
    n = 20_000_000
    names = arrays.bin_names()
    row = backend.trace_lights(arrays, [arrays.light], n, seed=61, device=0)[0]
    ref = _oracle_row(arrays, native, n, seed=62)
    assert row[:len(names)].sum() == n
    assert not bins_outside_band(names, row[:len(names)], ref[:len(names)], z=4.5)
    edges = sum(int(c) for name, c in zip(names, row) if name.startswith("exit/plate/") and name[-2:] in ("-x", "+x", "-y", "+y"))
    assert edges > 0.15 * n                                 # the case is about trapped light (18.6 % leaves through the edges)
    # events per photon have a long tail (trapped photons bounce up to 1000 times): sigma of the two-sample ratio ~4e-4
    assert abs(row[len(names)] / ref[len(names)] - 1.0) < 1.5e-3, (row[len(names)] / n, ref[len(names)] / n)
