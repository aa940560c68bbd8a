"""Diagnostic: tallies of two library builds side by side (grouped bins), same seeds.  python scripts/compare_libs.py libA[:ENV=1] libB ..."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200
lscpm_b200.install()
from lscpm_b200 import backend, flatten
from helpers import standard_scene, am15_like_photon_spectrum, grouped
arrays = flatten.flatten_scene(standard_scene(tilt=40, elevation=35, azimuth=150, spectrum=am15_like_photon_spectrum()))
rows = backend.trace_lights(arrays, [arrays.light] * 32, 4_000_000, seed=77)
names = arrays.bin_names()
tot = rows[:, :len(names)].sum(axis=0)
g = grouped(names, tot)
out = {k: int(v) for k, v in g.items()}
out["events"] = int(rows[:, len(names)].sum()); out["emissions"] = int(rows[:, len(names) + 1].sum())
print(json.dumps(out))
'''
res = []
for spec in sys.argv[1:]:
    lib, *envs = spec.split(":")
    env = dict(os.environ, LSCPM_LIBRARY=os.path.abspath(lib), **dict(e.split("=") for e in envs))
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True)
    if not r.stdout.strip():
        print(r.stderr[-800:]); sys.exit(1)
    res.append(json.loads(r.stdout.strip().splitlines()[-1]))
keys = sorted(set().union(*[r.keys() for r in res]))
n = 32 * 4_000_000
print(f"{'bin family':40s}" + "".join(f"{os.path.basename(s)[:22]:>24s}" for s in sys.argv[1:]))
for k in keys:
    vals = [r.get(k, 0) for r in res]
    sig = (vals[0] - vals[-1]) / max((vals[0] + vals[-1]) ** 0.5, 1)
    print(f"{k:40s}" + "".join(f"{v / n:24.6f}" for v in vals) + f"   z(first-last) {sig:+.1f}")
