"""Trace one of compare_kernels.py's scenes with the kernel the library picks (or LSCPM_TRACE_KERNEL), a given number of
times: the command behind the per-scene ncu captures and the compute-sanitizer runs.

    python scripts/trace_case.py pv_both 1000000 64 3
    compute-sanitizer --tool racecheck python scripts/trace_case.py am15 20000 3 1
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import lscpm_b200
lscpm_b200.install()
from lscpm_b200 import backend, flatten
from helpers import standard_scene, am15_like_photon_spectrum

CASES = {
    "585nm": dict(tilt=0, elevation=90, wavelength=585.0),
    "am15": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()),
    "diffuse": dict(tilt=30, diffuse=True, spectrum=am15_like_photon_spectrum()),
    "no_dye": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum(), include_dye=False),
    "pv_both": dict(tilt=0, elevation=60, spectrum=am15_like_photon_spectrum(), add_bottom_PV=True, add_side_PV=True),
    "pv_bottom_tilted": dict(tilt=40, elevation=50, wavelength=555.0, add_bottom_PV=True),
    "pv_no_dye": dict(tilt=10, elevation=60, spectrum=am15_like_photon_spectrum(), include_dye=False, add_bottom_PV=True, add_side_PV=True),
}

label = sys.argv[1] if len(sys.argv) > 1 else "am15"
ppb = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 64
repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 3
arrays = flatten.flatten_scene(standard_scene(**CASES[label]))
scene = backend.SceneHandle(arrays, 0)
plan = backend.Plan(scene, [arrays.light] * nb, list(range(nb)))
tallies = plan.new_tallies()
for r in range(repeats):
    tallies.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.trace_into(tallies, ppb, 11); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    row = tallies.sum(dim=0).cpu().numpy()
    assert row[:scene.n_bins].sum() == nb * ppb
    print(f"{label}: {scene.kernel_name}: {nb} x {ppb} photons in {ms:.2f} ms ({nb * ppb / ms / 1e3:.0f} Mph/s), "
          f"events/photon {row[scene.n_bins] / (nb * ppb):.4f}", flush=True)
