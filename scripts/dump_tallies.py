"""Dump the tallies of a few cases to an .npz (to compare two builds of the library: same seed -> same tallies)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import lscpm_b200
lscpm_b200.install()
from lscpm_b200 import backend, flatten
from helpers import standard_scene, am15_like_photon_spectrum

CASES = {
    "am15": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()),
    "tilt40_oblique": dict(tilt=40, elevation=35, azimuth=140, spectrum=am15_like_photon_spectrum()),
    "diffuse": dict(tilt=30, diffuse=True, spectrum=am15_like_photon_spectrum()),
    "no_dye": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum(), include_dye=False),
    "pv_both": dict(tilt=0, elevation=60, spectrum=am15_like_photon_spectrum(), add_bottom_PV=True, add_side_PV=True),
}
out = {}
for label, kw in CASES.items():
    arrays = flatten.flatten_scene(standard_scene(**kw))
    scene = backend.SceneHandle(arrays, 0)
    plan = backend.Plan(scene, [arrays.light] * 8, list(range(8)))
    tallies = plan.new_tallies()
    plan.trace_into(tallies, 500_000, 99); torch.cuda.synchronize()
    tallies.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.trace_into(tallies, 4_000_000, 99); e1.record(); torch.cuda.synchronize()
    out[label] = tallies.cpu().numpy()
    print(f"{label}: {e0.elapsed_time(e1):.2f} ms", flush=True)
np.savez(sys.argv[1], **out)
