import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
import c_oracle
from lscpm_b200 import _native as nat, backend, flatten
from lscpm_b200.workloads import standard_scene
from helpers import grouped
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
for wl in (650.0, 720.0, 760.0, 800.0):
  for dye in (False,):
    arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=25, azimuth=100, wavelength=wl, include_dye=dye))
    names = arrays.bin_names()
    gpu = backend.trace_lights(arrays, [arrays.light], n, seed=99, device=0)[0]
    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    b, _ = backend.pack_batches([arrays.light])
    ref = oracle.trace(b[0], photon_count=n, seed=100)
    g, r = grouped(names, gpu[:len(names)]), grouped(names, ref[:len(names)])
    print(wl, "dye", dye, "events/photon", gpu[len(names)] / n, ref[len(names)] / n)
    for k in sorted(set(g) | set(r)):
        a, c = g.get(k, 0), r.get(k, 0)
        if a or c: print(f"   {k:28s} gpu {a/n:.6f} oracle {c/n:.6f}  z = {(a - c) / np.sqrt(max(a + c, 1)):+.2f}")
    for nm, a, c in zip(names, gpu, ref):
        if (a or c) and nm.startswith("exit/LSC") : print(f"      {nm:40s} gpu {a/n:.6f} oracle {c/n:.6f}  z = {(a - c) / np.sqrt(max(a + c, 1)):+.2f}")
