"""Row f1/f2 of the scope table on the CPU: solar position against the ephemerides committed in the reference's result
CSVs, the reference's own solar_data row-count test, light descriptors against flattened scenes, and the batched
drivers' logic and CSV schema with the oracle injected as the tracer."""
import os

import numpy as np
import pandas as pd
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sites():
    from miniplant.locations import LOCATIONS

    return {site.name: site for site in LOCATIONS}


def test_solar_position_matches_reference_ephemerides():
    gold = pd.read_csv(os.path.join(GOLDEN, "reference_ephemerides.csv"), index_col=0)
    sites = _sites()
    for name, group in gold.groupby("site"):
        times = pd.to_datetime(group.index, utc=True)
        pos = sites[name].get_solarposition(times)
        d_el = np.abs(pos["apparent_elevation"].values - group["apparent_elevation"].values)
        d_az = np.abs(((pos["azimuth"].values - group["azimuth"].values + 180) % 360) - 180)
        # azimuth is ill-conditioned near the zenith (Townsville): weight by cos(elevation)
        d_az = d_az * np.cos(np.radians(group["apparent_elevation"].values))
        assert d_el.max() < 0.02, (name, d_el.max())
        assert d_az.max() < 0.03, (name, d_az.max())


def test_solar_data_for_place_and_time():
    """reference tests/test_solar_data.py:5-23, unchanged in substance."""
    from miniplant.solar_data import solar_data_for_place_and_time

    tilt = {"Townsville": -10, "Eindhoven": 40, "Plataforma Solar de Almería": 30, "North Cape": 50}
    points = {"Townsville": 366, "Eindhoven": 366, "Plataforma Solar de Almería": 366, "North Cape": 298}
    for name, site in _sites().items():
        frame = solar_data_for_place_and_time(site, tilt[name], 60 * 60 * 12, write_csv=False)
        assert frame["azimuth"].count() == points[name]


def test_spectrl2_reproduces_the_reference_irradiances():
    """Pins the SPECTRL2 restatement on reference output: the direct / sky-diffuse irradiances pvlib produced for the
    reference's committed full-year runs (recovered from the result CSVs, tests/golden/make_ephemerides.py), 1394 time
    points over four sites and five tilts, fed with the CSVs' own solar positions.  Relative error <= 1e-4 everywhere
    (measured: median 2e-5, maximum 7e-5)."""
    from lscpm_b200.compat import pvlib_lite
    from miniplant import solar_data
    from miniplant.utils import spectral_distribution_to_photon_distribution
    from lscpm_b200.compat.pvtrace_api import Distribution

    gold = pd.read_csv(os.path.join(GOLDEN, "reference_irradiances.csv"), index_col=0)
    sites = _sites()
    for (name, tilt), group in gold.groupby(["site", "tilt"]):
        site = sites[name]
        doy = pd.to_datetime(group.index, utc=True).tz_convert(site.tz).dayofyear.values
        zenith = 90.0 - group["apparent_elevation"].values
        spec = pvlib_lite.spectrum.spectrl2(
            apparent_zenith=zenith, aoi=pvlib_lite.irradiance.aoi(tilt, 180, zenith, group["azimuth"].values),
            surface_tilt=tilt, ground_albedo=solar_data.albedo, surface_pressure=pvlib_lite.alt2pres(site.altitude),
            relative_airmass=pvlib_lite.relative_airmass(zenith), precipitable_water=solar_data.water_vapor_content,
            ozone=solar_data.ozone, aerosol_turbidity_500nm=solar_data.tau500, dayofyear=doy)
        assert spec["wavelength"].shape == (27,) and spec["poa_direct"].shape == (27, len(group))
        for key, column in (("poa_direct", "direct_irradiance"), ("poa_sky_diffuse", "diffuse_irradiance")):
            got = np.empty(len(group))
            for i in range(len(group)):   # the reference's own conversion and integration (utils.py:26-36, solar_data.py:114-119)
                d = spectral_distribution_to_photon_distribution(Distribution(spec["wavelength"], spec[key][:, i]), 1800)
                got[i] = np.trapezoid(d._y, d._x)
            err = np.abs(got / group[column].values - 1)
            assert err.max() < 1e-4, (name, tilt, key, err.max())


def test_year_has_the_reference_row_count_and_irradiance():
    """The whole solar stage (own solar position -> air mass -> SPECTRL2 restatement): same 8048 half-hour points as
    the reference's Eindhoven 40 deg run, irradiances of the experimental window within 1e-3, yearly irradiance-weighted
    mean within 2e-4 of the reference's (a 0.015 deg position error only matters at sunrise and sunset)."""
    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_for_place_and_time

    frame = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
    assert len(frame) == 8048
    exp = pd.read_csv(os.path.join(GOLDEN, "reference_experimental_conditions.csv"), index_col=0)
    window = frame.loc["2020-07-15 15:00":"2020-07-15 17:30"]
    assert len(window) == len(exp) == 6
    ref_direct = (exp["direct_reacted"] / exp["simulation_direct"] / 0.47 ** 2).values
    ref_diffuse = (exp["diffuse_reacted"] / exp["simulation_diffuse"] / 0.47 ** 2).values
    assert np.allclose(window["direct_irradiance"].values, ref_direct, rtol=1e-3)
    assert np.allclose(window["diffuse_irradiance"].values, ref_diffuse, rtol=1e-3)
    assert np.allclose(window["apparent_elevation"].values, exp["apparent_elevation"].values, atol=0.01)
    spectrum = window["direct_spectrum"].iloc[0]
    assert len(spectrum._x) == 27 and spectrum._x[0] == 360 and spectrum._x[-1] == 690
    gold = pd.read_csv(os.path.join(GOLDEN, "reference_irradiances.csv"), index_col=0)
    gold = gold[(gold.site == "Eindhoven") & (gold.tilt == 40)]
    ours = frame.copy()
    ours.index = ours.index.map(str)
    ours = ours.loc[gold.index]
    for column in ("direct_irradiance", "diffuse_irradiance"):
        ratio = ours[column].values / gold[column].values
        assert abs(np.median(ratio) - 1) < 2e-4, (column, np.median(ratio))
        assert abs(np.average(ratio, weights=gold[column].values) - 1) < 2e-4, column


@pytest.mark.parametrize("tilt,elev,az", [(0, 90, 180), (40, 50, 180), (-10, 35.5, 213.5), (25, 5, 95)])
def test_light_helpers_equal_flattened_scenes(tilt, elev, az):
    from lscpm_b200 import flatten, lights
    from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene

    spec = am15_like_photon_spectrum()
    a = flatten.flatten_light(standard_scene(tilt=tilt, elevation=elev, azimuth=az, spectrum=spec))
    b = lights.direct_light(tilt, elev, az, spec)
    c = flatten.flatten_light(standard_scene(tilt=tilt, diffuse=True, wavelength=610.0))
    d = lights.diffuse_light(tilt, None, 610.0)
    for x, y in ((a, b), (c, d)):
        assert x.kind == y.kind and x.back_off == pytest.approx(y.back_off) and x.tilt_deg == y.tilt_deg
        assert np.allclose(x.mask_pose, y.mask_pose, atol=1e-12) and np.allclose(x.mask_half, y.mask_half)
        assert np.allclose(x.direction, y.direction, atol=1e-12)
        assert x.wavelength_nm == y.wavelength_nm
        assert (x.spectrum is None) == (y.spectrum is None)
        if x.spectrum is not None:
            assert np.array_equal(x.spectrum, y.spectrum)


def _oracle_tracer(arrays, light_list, photons, seed):
    import c_oracle
    from lscpm_b200 import _native as nat, backend

    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    rows = []
    for i, light in enumerate(light_list):
        batches, _ = backend.pack_batches([light], [i])
        spectra = [light.spectrum] if light.spectrum is not None else None
        rows.append(oracle.trace(batches[0], spectra=spectra, photon_count=photons, seed=seed))
    return np.stack(rows)


def test_yearlong_simulation_schema_and_values(tmp_path):
    """Driver logic with the oracle as tracer: same CSV columns as reference full_simulation.py:115-125."""
    from miniplant.full_simulation import yearlong_simulation
    from miniplant.locations import EINDHOVEN
    from miniplant.scene_creator import REACTOR_AREA_IN_M2
    from miniplant.solar_data import solar_data_for_place_and_time

    data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
    target = tmp_path / "out" / "result.csv"
    start, end = pd.Timestamp("2020-07-15 15:00", tz=EINDHOVEN.tz), pd.Timestamp("2020-07-15 17:30", tz=EINDHOVEN.tz)
    res = yearlong_simulation(40, EINDHOVEN, workers=12, num_photons_per_simulation=3000, time_range=(start, end),
                              target_file=target, solar_data=data, seed=5, tracer=_oracle_tracer)
    written = pd.read_csv(target, index_col=0)
    assert list(written.columns) == ["apparent_elevation", "azimuth", "simulation_direct", "direct_reacted",
                                     "simulation_diffuse", "diffuse_reacted"]
    assert len(written) == 6
    assert np.allclose(written["direct_reacted"], written["simulation_direct"] * res["direct_irradiance"].values * REACTOR_AREA_IN_M2)
    # statistical anchor: the reference measured 0.216-0.225 (direct) and 0.209-0.225 (diffuse) at 1e4 photons/point
    assert 0.18 < written["simulation_direct"].mean() < 0.26
    assert 0.17 < written["simulation_diffuse"].mean() < 0.26


def test_evaluate_tilt_angle_schema_and_skip(tmp_path):
    from lscpm_b200.compat.pvtrace_api import Distribution
    from miniplant import angle_optimization_simulation as aos
    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_from_positions

    data = solar_data_from_positions([30.0, 45.0, 60.0], [150.0, 180.0, 210.0], 30, index=pd.RangeIndex(3), direct_irradiance=2.0)
    zero = Distribution(np.array([360.0, 690.0]), np.zeros(2))
    data.at[1, "direct_spectrum"] = zero                       # sunrise/sunset point: skipped (reference :65-67)
    target = tmp_path / "tilt.csv"
    res = aos.evaluate_tilt_angle(30, EINDHOVEN, workers=12, solar_data=data, seed=1, tracer=_oracle_tracer, target_file=target)
    written = pd.read_csv(target, index_col=0)
    assert list(written.columns) == ["apparent_elevation", "azimuth", "simulation_direct", "direct_reacted"]
    assert written["simulation_direct"].iloc[1] == 0 and written["direct_reacted"].iloc[1] == 0
    assert all(0.05 < v < 0.45 for v in written["simulation_direct"].iloc[[0, 2]])
    assert np.allclose(written["direct_reacted"], written["simulation_direct"] * 2.0)
    assert aos.RAYS_PER_SIMULATIONS == 100 and aos.INCLUDE_DYE is True


def test_oracle_against_the_reference_yearly_results():
    """Pins the ORACLE on reference output: the per-time-point counts of the full-year runs the reference committed
    (tests/golden/reference_yearly_counts.npz, 120 photons per point).  Every 16th time point of the Eindhoven 40 deg
    year, 1500 photons each, traced by the C oracle and compared on the SAME rows: the difference of the two summed
    fractions in units of its binomial standard deviation (both sides' noise) stays within 4."""
    from helpers import frame_epoch, reference_counts
    from lscpm_b200 import lights
    from miniplant.full_simulation import trace_time_points
    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_for_place_and_time

    gold = np.load(os.path.join(GOLDEN, "reference_yearly_counts.npz"))
    data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
    rows = data.iloc[3::16]
    n_ours = 1500
    for dye, key in ((True, "Eindhoven_40deg_results"), (False, "Eindhoven_40deg_results_no_dye")):
        ref = reference_counts(gold, key)
        assert len(data) == len(ref)
        direct = [lights.direct_light(40, r["apparent_elevation"], r["azimuth"], r["direct_spectrum"]) for _, r in rows.iterrows()]
        diffuse = [lights.diffuse_light(40, r["diffuse_spectrum"]) for _, r in rows.iterrows()]
        frac = trace_time_points(40, direct + diffuse, n_ours, include_dye=dye, seed=4, tracer=_oracle_tracer)
        n = len(rows)
        ours = pd.DataFrame({"direct": frac[:n], "diffuse": frac[n:]}, index=frame_epoch(rows))
        both = ours.join(ref, how="inner", rsuffix="_ref")
        assert len(both) == n
        for column in ("direct", "diffuse"):
            p_ours, p_ref = both[column].values, both[column + "_ref"].values / 120.0
            pooled = 0.5 * (p_ours + p_ref)
            sigma = np.sqrt(np.sum(pooled * (1 - pooled)) * (1 / 120.0 + 1 / n_ours))
            z = (p_ours.sum() - p_ref.sum()) / sigma
            assert abs(z) < 4.0, (dye, column, z, p_ours.mean(), p_ref.mean())


def test_experimental_conditions_window():
    """The 15 July 2020 15:00-17:30 window of the reference's 1e4-photon run (its experimental_comparison script builds it
    with ``Location.pytz.localize``): the solar stage yields exactly the six half-hour points of the committed result CSV."""
    import datetime

    from miniplant.locations import EINDHOVEN
    from miniplant.solar_data import solar_data_for_place_and_time

    start = EINDHOVEN.pytz.localize(datetime.datetime(2020, 7, 15, 15, 0))
    end = EINDHOVEN.pytz.localize(datetime.datetime(2020, 7, 15, 17, 30))
    assert start.utcoffset().total_seconds() == 7200
    frame = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
    window = frame.loc[start: end]
    exp = pd.read_csv(os.path.join(GOLDEN, "reference_experimental_conditions.csv"), index_col=0)
    assert [str(t) for t in window.index] == list(exp.index)
