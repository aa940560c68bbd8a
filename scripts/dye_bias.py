"""Diagnostic (not product): how far do alternatives to the restated luminophore rules move the reacted fraction,
measured against the reference's committed tilt-sweep counts (tests/golden/reference_sweep_counts.npz)?
Needs the -DLSCPM_DIAG build of the library (variants/diag.so).  Usage: python scripts/dye_bias.py [photons_per_point]"""
import os, pickle, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
N = int(float(sys.argv[1])) if len(sys.argv) > 1 and sys.argv[1] != "--child" else 20000
ALL = os.environ.get("DIAG_ALL_FILES")   # every committed dye sweep instead of the 14-file subset
FILES = ["Eindhoven_0deg", "Eindhoven_20deg", "Eindhoven_40deg", "Eindhoven_60deg", "Eindhoven_80deg", "Plataforma Solar de Almería_15deg",
         "Plataforma Solar de Almería_30deg", "Plataforma Solar de Almería_45deg", "Townsville_0deg", "Townsville_15deg", "Townsville_30deg",
         "North Cape_40deg", "North Cape_55deg", "North Cape_70deg"]
CACHE = "/tmp/dye_bias_inputs.pkl"
VARIANTS = [("base", {}), ("redshift", {"LSCPM_DIAG_EMIT": "1"}), ("full", {"LSCPM_DIAG_EMIT": "2"}), ("3kT", {"LSCPM_DIAG_EMIT": "3"}),
            ("qy0.94", {"DIAG_QY": "0.94"}), ("ems_trapz", {"DIAG_EMS_TRAPZ": "1"}), ("pmma_bg0.3", {"DIAG_BG": "0.3"})]
if os.path.exists(os.path.join(ROOT, "variants", "diag_old.so")):   # the library before the adjacency rule was corrected
    VARIANTS += [("old_rules", {"DIAG_LIB": "diag_old.so"}), ("old_no_sibling", {"DIAG_LIB": "diag_old.so", "LSCPM_DIAG_NO_SIBLING": "1"})]

if "--child" in sys.argv:
    import numpy as np, pandas as pd
    import lscpm_b200; lscpm_b200.install()
    from helpers import binomial_z, reference_counts
    from lscpm_b200 import backend, flatten
    from miniplant.scene_creator import create_direct_scene
    n = int(sys.argv[sys.argv.index("--child") + 1])
    inputs = pickle.load(open(CACHE, "rb"))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_sweep_counts.npz"))
    k_sum = e_sum = v_sum = 0.0
    for key, (tilt, light_list, epoch) in inputs.items():
        arrays = flatten.flatten_scene(create_direct_scene(tilt_angle=tilt, include_dye=True), with_light=False)
        if os.environ.get("DIAG_QY"):
            arrays.comp_quantum_yield = np.where(arrays.comp_quantum_yield > 0, float(os.environ["DIAG_QY"]), arrays.comp_quantum_yield)
        if os.environ.get("DIAG_BG"):
            first = int(arrays.mat_comp_begin[int(arrays.node_material[1])])      # the plate's background absorber
            arrays.comp_coefficient = arrays.comp_coefficient.copy(); arrays.comp_coefficient[first] = float(os.environ["DIAG_BG"])
        if os.environ.get("DIAG_EMS_TRAPZ"):   # cumsum of the interval means = trapezoid CDF of the emission table
            t = int(arrays.comp_ems_table[arrays.comp_ems_table >= 0][0])
            lo, hi = int(arrays.table_begin[t]), int(arrays.table_begin[t + 1])
            y = arrays.table_y.copy(); seg = y[lo:hi].copy()
            y[lo] = 0.0; y[lo + 1:hi] = 0.5 * (seg[1:] + seg[:-1]); arrays.table_y = y
        rows = backend.trace_lights(arrays, light_list, n, seed=3)
        names = arrays.bin_names()
        react = np.array([name.startswith("react/") for name in names])
        p = rows[:, :len(names)][:, react].sum(axis=1) / float(n)
        both = pd.DataFrame({"p": p}, index=epoch).join(reference_counts(gold, key), how="inner")
        k_sum += both.direct.sum(); e_sum += 100 * both.p.sum(); v_sum += 100 * (both.p * (1 - both.p)).sum()
    print(f"ours/ref {e_sum / k_sum:.5f}  pooled z {(k_sum - e_sum) / np.sqrt(v_sum):+.2f}  (reference photons {100 * sum(len(v[2]) for v in inputs.values())})")
    sys.exit(0)

import lscpm_b200; lscpm_b200.install()
from helpers import frame_epoch
from lscpm_b200 import lights
from miniplant.locations import LOCATIONS
from miniplant.solar_data import solar_data_for_place_and_time
sites = {s.name: s for s in LOCATIONS}
if ALL:
    import numpy as _np
    _g = _np.load(os.path.join(ROOT, "tests", "golden", "reference_sweep_counts.npz"))
    FILES = sorted({n.split("|")[0] for n in _g.files if "_no_dye_" not in n})
    VARIANTS = VARIANTS[:1]
inputs = {}
for key in FILES:
    site = sites[key.split("_")[0]]
    tilt = int(key.split("_")[-1].replace("deg", "")) * (-1 if site.name == "Townsville" else 1)
    data = solar_data_for_place_and_time(site, tilt, 1800, write_csv=False)
    inputs[key] = (tilt, [lights.direct_light(tilt, r["apparent_elevation"], r["azimuth"], r["direct_spectrum"]) for _, r in data.iterrows()],
                   frame_epoch(data))
pickle.dump(inputs, open(CACHE, "wb"))
for name, env in VARIANTS:
    e = dict(os.environ, **env)
    e["LSCPM_LIBRARY"] = os.path.join(ROOT, "variants", e.pop("DIAG_LIB", "diag.so"))
    r = subprocess.run([sys.executable, __file__, "--child", str(N)], env=e, capture_output=True, text=True)
    print(f"{name:12s} {r.stdout.strip() or r.stderr.strip()[-600:]}", flush=True)
