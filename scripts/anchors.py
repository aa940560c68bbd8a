"""The mirrored drivers against everything the reference committed from its own pvtrace runs (for DESIGN.md section 5).

For each of the ten full-year result CSVs (120 photons per time point and light source, tests/golden/
reference_yearly_counts.npz) the whole chain runs here - own solar position, SPECTRL2 restatement, scene, CUDA tracing at
N photons per point - and the reference's per-point counts k_i ~ Binomial(120, p_i) are tested against our p_i:

    z = (sum k_i - 120 sum p_i) / sqrt(120 sum p_i (1 - p_i))

over the whole year and inside elevation bands (the angular dependence), plus the irradiance-weighted yearly yield.
Usage: python scripts/anchors.py [photons_per_point=20000]
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pandas as pd

import lscpm_b200

lscpm_b200.install()
from helpers import binomial_z, frame_epoch, reference_counts
from lscpm_b200 import lights
from miniplant.full_simulation import trace_time_points, year_batches, yearlong_simulation
from miniplant.locations import LOCATIONS
from miniplant.solar_data import solar_data_for_place_and_time

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_yearly_counts.npz"))
SITES = {s.name: s for s in LOCATIONS}
BANDS = [(0, 10), (10, 20), (20, 35), (35, 50), (50, 91)]
tmp = tempfile.mkdtemp()


def z_score(k, p, n_ref=120):
    var = n_ref * np.sum(p * (1 - p))
    return (k.sum() - n_ref * p.sum()) / np.sqrt(var) if var > 0 else 0.0


# the 1e4-photon run of 15 Jul 2020 (experimental_comparison/experimental_conditions_results.csv)
exp = pd.read_csv(os.path.join(ROOT, "tests", "golden", "reference_experimental_conditions.csv"), index_col=0)
site = SITES["Eindhoven"]
data = solar_data_for_place_and_time(site, 40, 1800, write_csv=False)
start, end = pd.Timestamp("2020-07-15 15:00", tz=site.tz), pd.Timestamp("2020-07-15 17:30", tz=site.tz)
r = yearlong_simulation(40, site, num_photons_per_simulation=4_000_000, time_range=(start, end),
                        target_file=os.path.join(tmp, "w.csv"), solar_data=data, seed=1)
for col in ("simulation_direct", "simulation_diffuse"):
    p = r[col].values
    k = exp[col].values * 1e4
    print(f"window {col}: ours {np.round(p, 4)} mean {p.mean():.4f} | reference {exp[col].values} mean {exp[col].mean():.4f}"
          f" | z = {z_score(k, p, 1e4):+.2f}")
for col, irr in (("direct_reacted", "direct_irradiance"), ("diffuse_reacted", "diffuse_irradiance")):
    print(f"window {col}: ours/reference {np.round(r[col].values / exp[col].values, 4)}")

all_z = []
for key in sorted({name.split("|")[0] for name in GOLD.files}):
    site_name, tilt_s = key.split("_")[0], key.split("_")[1]
    tilt = int(tilt_s.replace("deg", ""))
    dye = not key.endswith("no_dye")
    site = SITES[site_name]
    data = solar_data_for_place_and_time(site, tilt, 1800, write_csv=False)
    t0 = time.time()
    res = yearlong_simulation(tilt, site, num_photons_per_simulation=N, include_dye=dye, target_file=os.path.join(tmp, "y.csv"),
                              solar_data=data, seed=2)
    dt = time.time() - t0
    ours = pd.DataFrame({"pd": res["simulation_direct"].values, "ps": res["simulation_diffuse"].values,
                         "el": res["apparent_elevation"].values, "id": res["direct_irradiance"].values,
                         "is": res["diffuse_irradiance"].values}, index=frame_epoch(res))
    ref = reference_counts(GOLD, key).rename(columns={"direct": "kd", "diffuse": "ks"})
    both = ours.join(ref, how="inner")
    zd, zs = binomial_z(both.kd.values, both.pd.values, 120), binomial_z(both.ks.values, both.ps.values, 120)
    all_z += [zd, zs]
    bands = []
    for lo, hi in BANDS:
        m = (both.el.values >= lo) & (both.el.values < hi)
        if m.sum() > 20:
            z = binomial_z(both.kd.values[m], both.pd.values[m], 120)
            all_z.append(z)
            bands.append(f"{lo}-{hi}:{z:+.1f}")
    # irradiance-weighted yearly yield (what the reference's analysis scripts sum): ours / reference
    y_ref = (both.kd / 120 * both.id).sum() + (both.ks / 120 * both["is"]).sum()
    y_our = (both.pd * both.id).sum() + (both.ps * both["is"]).sum()
    print(f"{key:52s} rows {len(res)}/{len(ref)} matched {len(both)} | mean direct {both.pd.mean():.4f} vs {both.kd.mean() / 120:.4f} z {zd:+.2f}"
          f" | diffuse {both.ps.mean():.4f} vs {both.ks.mean() / 120:.4f} z {zs:+.2f} | direct z by elevation band {' '.join(bands)}"
          f" | yearly yield ours/ref {y_our / y_ref:.4f} | traced {2 * len(res) * N / 1e6:.0f} Mphotons in {dt:.1f}s")
all_z = np.array(all_z)
print(f"{len(all_z)} z-scores: mean {all_z.mean():+.2f}, rms {np.sqrt((all_z ** 2).mean()):.2f}, max |z| {np.abs(all_z).max():.2f}")


# ---- the committed tilt sweeps (angle_optimization_simulation.py: 100 photons per point, direct light only) ------------
SWEEP = np.load(os.path.join(ROOT, "tests", "golden", "reference_sweep_counts.npz"))
print("\ntilt sweeps (100 photons per point): file, rows, mean ours vs reference, z, yearly sum(direct_reacted) ours/ref")
sweep_z = {True: [], False: []}
pooled = {True: [0.0, 0.0, 0.0], False: [0.0, 0.0, 0.0]}
for key in sorted({name.split("|")[0] for name in SWEEP.files}):
    dye = "_no_dye_" not in key
    site = SITES[key.split("_")[0]]
    tilt = int(key.split("_")[-1].replace("deg", ""))
    if site.name == "Townsville":
        tilt = -tilt                      # swept over range(0, -40, -5), file names carry |tilt|
    data = solar_data_for_place_and_time(site, tilt, 1800, write_csv=False)
    p = trace_time_points(tilt, year_batches(tilt, data, diffuse=False), N, include_dye=dye, seed=3)
    both = pd.DataFrame({"p": p, "irr": data["direct_irradiance"].values}, index=frame_epoch(data)).join(reference_counts(SWEEP, key), how="inner")
    z = binomial_z(both.direct.values, both.p.values, 100)
    sweep_z[dye].append(z)
    pooled[dye][0] += both.direct.sum(); pooled[dye][1] += 100 * both.p.sum(); pooled[dye][2] += 100 * (both.p * (1 - both.p)).sum()
    print(f"{key:44s} rows {len(data)}/{len(SWEEP[key + '|direct'])} | {both.p.mean():.4f} vs {both.direct.mean() / 100:.4f} z {z:+.2f}"
          f" | yield {(both.p * both.irr).sum() / (both.direct / 100 * both.irr).sum():.4f}")
for dye in (True, False):
    zs = np.array(sweep_z[dye])
    k, e, v = pooled[dye]
    print(f"{'dye' if dye else 'no dye'}: {len(zs)} files, z mean {zs.mean():+.2f} rms {np.sqrt((zs ** 2).mean()):.2f} max |z| {np.abs(zs).max():.2f};"
          f" pooled reference photons {e / 1:.0f} expected vs {k:.0f} counted: ratio ours/ref {e / k:.5f}, pooled z {(k - e) / np.sqrt(v):+.2f}")
