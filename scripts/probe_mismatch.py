"""Diagnostic: classify disagreements between lscpm_probe and the oracle's next_hit on rays harvested from oracle histories."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
import c_oracle
from lscpm_b200 import _native as nat, backend, flatten
from lscpm_b200.workloads import standard_scene
arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=25, azimuth=100, wavelength=585.0, include_dye=True))
oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
b, _ = backend.pack_batches([arrays.light])
pos, direc, wl = oracle.emit(b[0], 150000, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
hist = oracle.follow(pos, direc, wl, seed=6, max_steps=48)
P, D = [], []
for i in range(len(pos)):
    k = min(hist["n_steps"][i], 48)
    for s in range(k - 1):
        if hist["event"][i, s] in (0, 1, 2, 4):
            P.append(hist["pos"][i, s]); D.append(hist["dir"][i, s])
P = np.array(P).astype(np.float32).astype(np.float64); D = np.array(D).astype(np.float32).astype(np.float64)
D /= np.linalg.norm(D, axis=1, keepdims=True)
print(len(P), "rays")
ref = oracle.probe(P, D)
gpu = backend.probe_rays(arrays, P, D, device=0)
names = arrays.node_name
def kind(i): 
    if i < 0: return "none"
    n = names[i]; return "world" if i == 0 else "plate" if n.startswith("LSC") else "pfa" if n.startswith("Cap") else "mix"
clear = ref["distance"] > 1e-6
bad = clear & ~((gpu["hit"] == ref["hit"]) & (gpu["container"] == ref["container"]) & (gpu["adjacent"] == ref["adjacent"]))
print("clear", clear.sum(), "mismatch", bad.sum())
cat = collections.Counter()
for i in np.where(bad)[0]:
    cat[(kind(ref["container"][i]), kind(ref["hit"][i]), kind(ref["adjacent"][i]), "|", kind(gpu["container"][i]), kind(gpu["hit"][i]), kind(gpu["adjacent"][i]))] += 1
for k, v in cat.most_common(20): print(v, k)
idx = np.where(bad)[0][:12]
for i in idx:
    print(i, P[i], D[i], "ref", ref["hit"][i], ref["container"][i], ref["adjacent"][i], ref["distance"][i], "gpu", gpu["hit"][i], gpu["container"][i], gpu["adjacent"][i], gpu["distance"][i])
ok = clear & ~bad & (ref["hit"] > 0)
rel = np.abs(gpu["distance"][ok] - ref["distance"][ok]) / ref["distance"][ok]
print("distance rel err quantiles", np.quantile(rel, [0.5, 0.9, 0.99, 0.999, 1.0]))
