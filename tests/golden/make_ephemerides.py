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
