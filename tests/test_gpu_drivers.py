"""The batched drivers on the GPU (row f1): all time points in one launch, same CSV schema, results against the
statistical anchors the reference committed (experimental_conditions_results.csv, 1e4 photons per point)."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_experimental_window_against_reference_csv(tmp_path, built_library):
    from miniplant.full_simulation import yearlong_simulation
    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_for_place_and_time

    data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
    start, end = pd.Timestamp("2020-07-15 15:00", tz=EINDHOVEN.tz), pd.Timestamp("2020-07-15 17:30", tz=EINDHOVEN.tz)
    target = tmp_path / "experimental_conditions_results.csv"
    yearlong_simulation(tilt_angle=40, location=EINDHOVEN, workers=12, time_resolution=60 * 30, include_dye=True,
                        num_photons_per_simulation=1_000_000, time_range=(start, end), target_file=target,
                        solar_data=data, seed=11)
    ours = pd.read_csv(target, index_col=0)
    ref = pd.read_csv(os.path.join(GOLDEN, "reference_experimental_conditions.csv"), index_col=0)
    assert list(ours.columns) == list(ref.columns) and len(ours) == len(ref) == 6
    # the reference's fractions carry +-0.004 of Monte-Carlo noise per point (1e4 photons), +-0.0017 on the mean of six;
    # the irradiances behind *_reacted are reproduced to 1e-3 (tests/test_solar_stage.py)
    assert abs(ours["simulation_direct"].mean() - ref["simulation_direct"].mean()) < 0.006
    assert abs(ours["simulation_diffuse"].mean() - ref["simulation_diffuse"].mean()) < 0.006
    assert np.allclose(ours["direct_reacted"], ref["direct_reacted"], rtol=0.06)
    assert np.allclose(ours["diffuse_reacted"], ref["diffuse_reacted"], rtol=0.08)


def test_full_year_in_one_launch(tmp_path, built_library):
    from lscpm_b200 import _native
    from miniplant.full_simulation import yearlong_simulation
    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_for_place_and_time

    data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
    lib = _native.load()
    before = lib.lscpm_launch_count()
    res = yearlong_simulation(40, EINDHOVEN, num_photons_per_simulation=120, target_file=tmp_path / "year.csv",
                              solar_data=data, seed=3)
    assert lib.lscpm_launch_count() - before == 1              # 16 096 batches, one kernel launch
    assert len(res) == 8048
    # the reference's committed year (120 photons x 8048 points): mean 0.1994 direct, 0.2133 diffuse
    # two independent 120-photon years: the means differ by sqrt(2) x 0.0004 of noise
    assert abs(res["simulation_direct"].mean() - 0.1994) < 0.003
    assert abs(res["simulation_diffuse"].mean() - 0.2133) < 0.003
    assert set(np.round(res["simulation_direct"] * 120).astype(int) - res["simulation_direct"] * 120 < 1e-9) == {True}


def test_tilt_sweep_driver(tmp_path, built_library):
    from miniplant import angle_optimization_simulation as aos
    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_from_positions

    gold = pd.read_csv(os.path.join(GOLDEN, "reference_ephemerides.csv"), index_col=0)
    eind = gold[gold["site"] == "Eindhoven"]
    totals = {}
    for tilt in (0, 40, 80):
        data = solar_data_from_positions(eind["apparent_elevation"].values, eind["azimuth"].values, tilt, index=eind.index)
        res = aos.evaluate_tilt_angle(tilt, EINDHOVEN, workers=12, solar_data=data, seed=7, target_file=tmp_path / f"{tilt}.csv")
        assert len(res) == len(eind)
        totals[tilt] = res["simulation_direct"].mean()
    assert all(0.12 < v < 0.30 for v in totals.values())
    assert totals[0] != totals[40] != totals[80]


@pytest.mark.parametrize("site_name,tilt,dye,csv", [
    ("Eindhoven", 0, True, "Eindhoven_0deg_results"),
    ("Eindhoven", 40, True, "Eindhoven_40deg_results"),
    ("Townsville", -10, True, "Townsville_-10deg_results"),
    ("North Cape", 50, True, "North Cape_50deg_results"),
    ("Plataforma Solar de Almería", 30, True, "Plataforma Solar de Almería_30deg_results"),
    ("Eindhoven", 40, False, "Eindhoven_40deg_results_no_dye"),
    ("North Cape", 50, False, "North Cape_50deg_results_no_dye"),
    ("Townsville", -10, False, "Townsville_-10deg_results_no_dye"),
])
def test_yearly_counts_against_the_reference_results(site_name, tilt, dye, csv, tmp_path, built_library):
    """The whole chain (solar position, SPECTRL2 restatement, scene, CUDA tracing) against the reference's own pvtrace
    output: the per-time-point counts of its committed full-year runs (120 photons per point, ~1e6 photons per file and
    light source; tests/golden/reference_yearly_counts.npz).  Our p_i come from 2e4 photons per point; the reference's
    counts must be Binomial(120, p_i): z of the yearly totals and of three elevation bands within +-4, yearly means
    within 0.6 % (dye) / 1.2 % (no dye).  Measured over the ten files (scripts/anchors.py): |z| <= 2.0 except the
    Eindhoven 45 degree file, which disagrees with the reference's own sweep for the same tilt; DESIGN.md section 5."""
    from helpers import binomial_z, frame_epoch, reference_counts
    from miniplant.full_simulation import yearlong_simulation
    from miniplant.locations import LOCATIONS
    from miniplant.solar_data import solar_data_for_place_and_time

    gold = np.load(os.path.join(GOLDEN, "reference_yearly_counts.npz"))
    site = next(s for s in LOCATIONS if s.name == site_name)
    data = solar_data_for_place_and_time(site, tilt, 1800, write_csv=False)
    res = yearlong_simulation(tilt, site, num_photons_per_simulation=20000, include_dye=dye, target_file=tmp_path / "y.csv",
                              solar_data=data, seed=9)
    ours = pd.DataFrame({"pd": res["simulation_direct"].values, "ps": res["simulation_diffuse"].values,
                         "el": res["apparent_elevation"].values}, index=frame_epoch(res))
    ref = reference_counts(gold, csv).rename(columns={"direct": "kd", "diffuse": "ks"})
    both = ours.join(ref, how="inner")
    assert len(both) >= len(ref) - 1 and abs(len(res) - len(ref)) <= 1
    tol = 0.006 if dye else 0.012      # the reference's yearly means carry 0.2 % (dye) / 0.4 % (no dye) of Monte-Carlo noise
    assert abs(both.pd.mean() / (both.kd.mean() / 120) - 1) < tol
    assert abs(both.ps.mean() / (both.ks.mean() / 120) - 1) < tol
    assert abs(binomial_z(both.kd.values, both.pd.values, 120)) < 4.0
    assert abs(binomial_z(both.ks.values, both.ps.values, 120)) < 4.0
    for lo, hi in ((0, 15), (15, 35), (35, 91)):
        m = (both.el.values >= lo) & (both.el.values < hi)
        if m.sum() > 200:
            assert abs(binomial_z(both.kd.values[m], both.pd.values[m], 120)) < 4.0, (lo, hi)


def test_tilt_sweep_counts_against_the_reference_results(built_library):
    """The reference's committed tilt sweeps (angle_optimization_simulation.py, 100 photons per point, direct light):
    Eindhoven at 0 / 30 / 60 / 90 degrees with dye and 60 / 90 degrees without.  Same binomial test as above."""
    from helpers import binomial_z, frame_epoch, reference_counts
    from lscpm_b200 import lights
    from miniplant.full_simulation import trace_time_points
    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_for_place_and_time

    gold = np.load(os.path.join(GOLDEN, "reference_sweep_counts.npz"))
    for key, tilt, dye in (("Eindhoven_0deg", 0, True), ("Eindhoven_30deg", 30, True), ("Eindhoven_60deg", 60, True),
                           ("Eindhoven_90deg", 90, True), ("Eindhoven_no_dye_60deg", 60, False),
                           ("Eindhoven_no_dye_90deg", 90, False)):
        data = solar_data_for_place_and_time(EINDHOVEN, tilt, 1800, write_csv=False)
        batch = [lights.direct_light(tilt, r["apparent_elevation"], r["azimuth"], r["direct_spectrum"]) for _, r in data.iterrows()]
        p = trace_time_points(tilt, batch, 20000, include_dye=dye, seed=13)      # the per-row (list of LightArrays) form on purpose
        both = pd.DataFrame({"p": p}, index=frame_epoch(data)).join(reference_counts(gold, key), how="inner")
        assert len(both) >= len(gold[key + "|direct"]) - 2, key
        assert abs(binomial_z(both.direct.values, both.p.values, 100)) < 4.0, key
        assert abs(both.p.mean() / (both.direct.mean() / 100) - 1) < (0.008 if dye else 0.015), key


# ---- all 55 committed tilt sweeps, resolved by elevation band x azimuth band ---------------------------
#
# PRE-REGISTERED RULE SET (written before the first run of this test; the thresholds follow from the statistics of
# independent unit normals, not from the outcome).  For every sweep file the reference committed
# (tests/golden/reference_sweep_counts.npz: 55 files, 100 pvtrace photons per time point, 3.5e7 photons) our p_i at
# 2e4 photons per point give z = (sum k_i - 100 sum p_i) / sqrt(100 sum p_i (1 - p_i)) over a set of time points:
#   R1  per file, all points:                                   |z| < 4
#   R2  per file and (elevation band x azimuth band) cell with >= 300 points (about 400 cells):  |z| < 4.5
#   R3  the cell z-scores together: |mean| < 0.30 and rms < 1.25 (unit normals: 0 +- 0.05 and 1 +- 0.04; the margins leave
#       room for the 2e-3 relative Monte-Carlo noise of our own p_i and for sub-permille model differences)
#   R4  pooled over the dye files / the no-dye files: ours / reference within 0.2 % / 0.6 %
#   R5  held-out protocol: R1 holds separately on the even-numbered and on the odd-numbered time points of every file.
#       The bookkeeping rules of DESIGN.md section 4 were selected in round 1 on ALL points of these files; any further
#       rule selection has to use the even points only and report the odd ones as the held-out half.
ELEVATION_BANDS = ((0, 15), (15, 35), (35, 91))
AZIMUTH_BANDS = ((0, 150), (150, 210), (210, 361))


def test_all_committed_tilt_sweeps_by_elevation_and_azimuth_band(built_library, capsys):
    from helpers import binomial_z, frame_epoch, reference_counts
    from miniplant.full_simulation import trace_time_points, year_batches
    from miniplant.locations import LOCATIONS
    from miniplant.solar_data import solar_data_for_place_and_time

    gold = np.load(os.path.join(GOLDEN, "reference_sweep_counts.npz"))
    sites = {s.name: s for s in LOCATIONS}
    cells, report = [], []
    pooled = {True: [0.0, 0.0], False: [0.0, 0.0]}
    for key in sorted({name.split("|")[0] for name in gold.files}):
        dye = "_no_dye_" not in key
        site = sites[key.split("_")[0]]
        tilt = int(key.split("_")[-1].replace("deg", ""))
        if site.name == "Townsville":
            tilt = -tilt                      # swept over range(0, -40, -5); the file names carry |tilt|
        data = solar_data_for_place_and_time(site, tilt, 1800, write_csv=False)
        p = trace_time_points(tilt, year_batches(tilt, data, diffuse=False), 20000, include_dye=dye, seed=17)
        both = pd.DataFrame({"p": p, "el": data["apparent_elevation"].values, "az": data["azimuth"].values},
                            index=frame_epoch(data)).join(reference_counts(gold, key), how="inner")
        assert len(both) >= len(gold[key + "|direct"]) - 2, key
        k, q = both.direct.values, both.p.values
        z_all = binomial_z(k, q, 100)
        assert abs(z_all) < 4.0, (key, "R1", z_all)
        for half in (0, 1):
            z_half = binomial_z(k[half::2], q[half::2], 100)
            assert abs(z_half) < 4.0, (key, "R5", half, z_half)
        pooled[dye][0] += k.sum()
        pooled[dye][1] += 100 * q.sum()
        for e_lo, e_hi in ELEVATION_BANDS:
            for a_lo, a_hi in AZIMUTH_BANDS:
                m = (both.el.values >= e_lo) & (both.el.values < e_hi) & (both.az.values >= a_lo) & (both.az.values < a_hi)
                if m.sum() >= 300:
                    z = binomial_z(k[m], q[m], 100)
                    cells.append(z)
                    assert abs(z) < 4.5, (key, "R2", (e_lo, e_hi), (a_lo, a_hi), z)
        report.append(f"{key}: z {z_all:+.2f}")
    cells = np.array(cells)
    summary = (f"{len(report)} files, {len(cells)} cells: mean {cells.mean():+.3f} rms {np.sqrt((cells ** 2).mean()):.3f} "
               f"max |z| {np.abs(cells).max():.2f}; pooled ours/reference dye {pooled[True][1] / pooled[True][0]:.5f} "
               f"no dye {pooled[False][1] / pooled[False][0]:.5f}")
    with capsys.disabled():
        print("\n[sweep bands] " + summary)
    assert len(report) == 55 and len(cells) > 300
    assert abs(cells.mean()) < 0.30 and np.sqrt((cells ** 2).mean()) < 1.25, ("R3", summary)
    assert abs(pooled[True][1] / pooled[True][0] - 1) < 0.002 and abs(pooled[False][1] / pooled[False][0] - 1) < 0.006, ("R4", summary)
