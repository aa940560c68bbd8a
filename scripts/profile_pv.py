"""One warm-up and one measured launch on the photovoltaic-variant scene (for ncu / timing builds)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import lscpm_b200
lscpm_b200.install()
from lscpm_b200 import backend, flatten
from helpers import standard_scene, am15_like_photon_spectrum

nb, ppb = 64, 1_000_000
arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=60, spectrum=am15_like_photon_spectrum(), add_bottom_PV=True, add_side_PV=True))
scene = backend.SceneHandle(arrays, 0)
plan = backend.Plan(scene, [arrays.light] * nb, list(range(nb)))
tallies = plan.new_tallies()
for seed in (1, 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.trace_into(tallies, ppb, seed); e1.record(); torch.cuda.synchronize()
    print(f"seed {seed}: {e0.elapsed_time(e1):.2f} ms  {nb * ppb / e0.elapsed_time(e1) / 1e3:.0f} Mphotons/s", flush=True)
