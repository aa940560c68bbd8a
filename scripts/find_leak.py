import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
from lscpm_b200 import backend, flatten
from lscpm_b200.workloads import standard_scene
arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, wavelength=585.0, add_side_PV=True))
names = arrays.bin_names()
target = [i for i, n in enumerate(names) if n.startswith("exit/LSC") and n.split("/")[-1] in ("-x", "+x", "-y", "+y")]
found = 0
for chunk in range(8):
    n = 4_000_000
    pos, direc, wl = backend.emit_rays(arrays, arrays.light, n, seed=11 + chunk, device=0)
    out = backend.trace_rays(arrays, pos, direc, wl, seed=77, device=0, photon_begin=chunk * n)
    idx = np.where(np.isin(out["final_bin"], target))[0]
    for i in idx[:3]:
        rec = backend.trace_rays(arrays, pos[i:i+1], direc[i:i+1], wl[i:i+1], seed=77, device=0, photon_begin=chunk * n + int(i), max_steps=200)
        k = int(rec["n_steps"][0])
        print("photon", chunk, i, "bin", names[rec["final_bin"][0]], "steps", k)
        for s in range(max(0, k - 6), min(k, 200)):
            print("   ", s, "ev", rec["event"][0, s], "node", arrays.node_name[rec["node"][0, s]] if rec["node"][0, s] >= 0 else None,
                  "pos", np.array2string(rec["pos"][0, s], precision=9), "dir", np.array2string(rec["dir"][0, s], precision=6))
        found += 1
    if found >= 6: break
print("found", found)
