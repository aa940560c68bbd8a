"""Extracts a small ephemeris fixture (timestamp, apparent_elevation, azimuth) from the result CSVs committed in the
reference (they are the only pvlib solar-position output available offline).  Run in the build container:

    python tests/golden/make_ephemerides.py
"""
import os

import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/miniplant/full_simulation_results"
SOURCES = {
    "Eindhoven": "Eindhoven/Eindhoven_40deg_results.csv",
    "Townsville": "Townsville/Townsville_-10deg_results.csv",
    "North Cape": "North Cape/North Cape_50deg_results.csv",
    "Plataforma Solar de Almería": "Plataforma Solar de Almería/Plataforma Solar de Almería_30deg_results.csv",
}
frames = []
for site, rel in SOURCES.items():
    df = pd.read_csv(os.path.join(REF, rel), index_col=0)[["apparent_elevation", "azimuth"]]
    df = df.iloc[::41].copy()          # ~200 rows per site, spread over the year and the day
    df.insert(0, "site", site)
    frames.append(df)
out = pd.concat(frames)
out.index.name = "timestamp"
out.to_csv(os.path.join(HERE, "reference_ephemerides.csv"))
exp = pd.read_csv("/root/reference/miniplant/experimental_comparison/experimental_conditions_results.csv", index_col=0)
exp.index.name = "timestamp"
exp.to_csv(os.path.join(HERE, "reference_experimental_conditions.csv"))
print(len(out), "ephemeris rows;", len(exp), "experimental rows")

# yearly column means of the committed full-year results (statistical anchors for the driver tests)
import glob
import json

means = {}
for path in sorted(glob.glob(os.path.join(REF, "*", "*.csv"))):
    frame = pd.read_csv(path, index_col=0)
    means[os.path.basename(path)] = {"rows": len(frame), "simulation_direct_mean": float(frame.simulation_direct.mean()),
                                     "simulation_diffuse_mean": float(frame.simulation_diffuse.mean())}
with open(os.path.join(HERE, "reference_yearly_means.json"), "w") as fh:
    json.dump({"source": "column means of the result CSVs committed under miniplant/full_simulation_results/ in the reference "
                         "(120 photons per time point)", "files": means}, fh, indent=1, ensure_ascii=False)

# irradiances the reference computed with pvlib's SPECTRL2 (solar_data.py:106-119), recovered from its result CSVs:
# direct_reacted = simulation_direct * direct_irradiance * area (full_simulation.py:86-88), likewise for the diffuse part
AREA = 0.47 * 0.47
IRRADIANCE_SOURCES = dict(SOURCES, **{"Eindhoven@0": "Eindhoven/Eindhoven_0deg_results.csv"})
frames = []
for key, rel in IRRADIANCE_SOURCES.items():
    site = key.split("@")[0]
    tilt = float(os.path.basename(rel).split("_")[1].replace("deg", ""))
    df = pd.read_csv(os.path.join(REF, rel), index_col=0)
    df = df[(df.simulation_direct > 0) & (df.simulation_diffuse > 0)].iloc[::29]
    frames.append(pd.DataFrame({"site": site, "tilt": tilt, "apparent_elevation": df.apparent_elevation,
                                "azimuth": df.azimuth,
                                "direct_irradiance": df.direct_reacted / df.simulation_direct / AREA,
                                "diffuse_irradiance": df.diffuse_reacted / df.simulation_diffuse / AREA}, index=df.index))
irr = pd.concat(frames)
irr.index.name = "timestamp"
irr.to_csv(os.path.join(HERE, "reference_irradiances.csv"))
print(len(irr), "irradiance rows")

# per-time-point photon counts of the committed full-year runs (120 photons per point and light source): the reference's
# own Monte-Carlo output, row by row, for the chi-square style comparison in tests/test_gpu_drivers.py
import numpy as np

EPOCH0 = 1577800800   # 2019-12-31T14:00Z, the earliest local midnight of 1 Jan 2020 among the sites; time points are
                      # stored as delta-coded half-hour slots after it (uint16)


def slots(index):
    epoch = np.asarray((pd.to_datetime(index, utc=True) - pd.Timestamp("1970-01-01", tz="UTC")) // pd.Timedelta(seconds=1), dtype=np.int64)
    assert np.all((epoch - EPOCH0) % 1800 == 0) and epoch.min() >= EPOCH0 and (epoch.max() - EPOCH0) // 1800 < 65536
    return np.diff((epoch - EPOCH0) // 1800, prepend=0).astype(np.uint16)   # delta-coded: cumsum() restores the slots


counts = {}
for path in sorted(glob.glob(os.path.join(REF, "*", "*.csv"))):
    frame = pd.read_csv(path, index_col=0)
    key = os.path.basename(path)[:-4]
    k_direct = np.rint(frame.simulation_direct.values * 120)
    k_diffuse = np.rint(frame.simulation_diffuse.values * 120)
    assert np.allclose(k_direct, frame.simulation_direct.values * 120, atol=1e-6) and k_direct.max() <= 120
    counts[key + "|slot"] = slots(frame.index)
    counts[key + "|direct"] = k_direct.astype(np.uint8)
    counts[key + "|diffuse"] = k_diffuse.astype(np.uint8)
np.savez_compressed(os.path.join(HERE, "reference_yearly_counts.npz"), **counts)
print(len(counts) // 3, "yearly count tables")

# per-time-point counts of the committed tilt-sweep runs (angle_optimization_simulation.py, 100 photons per point, direct
# light only): file name carries |tilt| (Townsville is swept over negative tilts, angle_optimization_simulation.py:116)
SWEEP = "/root/reference/miniplant/simulation_results"
sweep = {}
for site in ("Eindhoven", "North Cape", "Plataforma Solar de Almería", "Townsville"):
    for path in sorted(glob.glob(os.path.join(SWEEP, site, f"{site}_*deg_results.csv"))):
        frame = pd.read_csv(path, index_col=0)
        if "simulation_direct" not in frame.columns or len(frame) == 0:
            continue
        k = np.rint(frame.simulation_direct.values * 100)
        if not np.allclose(k, frame.simulation_direct.values * 100, atol=1e-6) or k.max() > 100:
            print("skipped (not 100 photons per point):", os.path.basename(path))
            continue
        key = os.path.basename(path)[:-len("_results.csv")]
        sweep[key + "|slot"] = slots(frame.index)
        sweep[key + "|direct"] = k.astype(np.uint8)
np.savez_compressed(os.path.join(HERE, "reference_sweep_counts.npz"), **sweep)
print(len(sweep) // 2, "tilt-sweep count tables")
