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
