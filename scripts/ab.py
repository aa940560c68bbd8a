"""A/B timing of library builds on the bench workload (config 2, 64 x 1e6 photons): median kernel time and a checksum
of the tallies (identical checksums = bit-identical results).  Usage: python scripts/ab.py lib1.so lib2.so ...
Each library runs in its own process (LSCPM_LIBRARY)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, hashlib
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import lscpm_b200
lscpm_b200.install()
from lscpm_b200 import backend, flatten
from helpers import standard_scene, am15_like_photon_spectrum
out = []
for label, kw, nb, ppb in (("am15", dict(spectrum=am15_like_photon_spectrum()), 64, 1_000_000), ("585", dict(wavelength=585.0), 64, 1_000_000),
                           ("nodye", dict(spectrum=am15_like_photon_spectrum(), include_dye=False), 64, 1_000_000)):
    arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, **kw))
    scene = backend.scene_handle(arrays, 0)
    plan = backend.Plan(scene, [arrays.light] * nb, list(range(nb)))
    tallies = plan.new_tallies()
    for w in range(2):
        plan.trace_into(tallies, ppb, 1)
    torch.cuda.synchronize()
    ms = []
    for k in range(7):
        tallies.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.trace_into(tallies, ppb, 2); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    digest = hashlib.md5(tallies.cpu().numpy().tobytes()).hexdigest()[:10]
    out.append(f"{label} {ms[len(ms)//2]:.3f} ms (min {ms[0]:.3f}) {digest}")
print(" | ".join(out))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ, LSCPM_LIBRARY=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):28s} {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
