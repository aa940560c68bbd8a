"""A/B of the trace kernels (lane-per-photon, block-sorted, pool configurations): tallies with the same seed and timing.

    python scripts/compare_kernels.py [case ...]      # default: all cases

The photon's fate depends only on (seed, batch id, photon index), so the two kernels must produce the same tallies up
to a handful of photons whose path flips on a last-bit difference in instruction contraction.
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import lscpm_b200
lscpm_b200.install()
from lscpm_b200 import backend, flatten
from helpers import standard_scene, am15_like_photon_spectrum

CASES = {
    "585nm": dict(tilt=0, elevation=90, wavelength=585.0),
    "am15": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()),
    "tilt40_oblique": dict(tilt=40, elevation=35, azimuth=140, spectrum=am15_like_photon_spectrum()),
    "diffuse": dict(tilt=30, diffuse=True, spectrum=am15_like_photon_spectrum()),
    "no_dye": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum(), include_dye=False),
    "pv_both": dict(tilt=0, elevation=60, spectrum=am15_like_photon_spectrum(), add_bottom_PV=True, add_side_PV=True),
    "pv_bottom_tilted": dict(tilt=40, elevation=50, wavelength=555.0, add_bottom_PV=True),
    "pv_no_dye": dict(tilt=10, elevation=60, spectrum=am15_like_photon_spectrum(), include_dye=False, add_bottom_PV=True, add_side_PV=True),
}


def handle(arrays, which):
    os.environ["LSCPM_TRACE_KERNEL"] = which
    return backend.SceneHandle(arrays, 0)


def timed(plan, tallies, ppb, seed):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.trace_into(tallies, ppb, seed); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


ONLY = sys.argv[1:]
for label, kw in CASES.items():
    if ONLY and label not in ONLY:
        continue
    arrays = flatten.flatten_scene(standard_scene(**kw))
    out = {}
    variants = ("lanes", "sorted", "pool120", "pool96", "pool64")
    for which in variants:
        scene = handle(arrays, which)
        for nb, ppb in ((3, 70_001), (64, 1_000_000)):
            plan = backend.Plan(scene, [arrays.light] * nb, list(range(nb)))
            tallies = plan.new_tallies()
            timed(plan, tallies, ppb, 7)
            tallies.zero_()
            ms = min(timed(plan, tallies, ppb, 11) for _ in range(2)); tallies.zero_(); timed(plan, tallies, ppb, 11)
            out[(which, nb)] = (tallies.cpu().numpy().copy(), ms)
    for nb, ppb, other in [(nb, ppb, o) for nb, ppb in ((3, 70_001), (64, 1_000_000)) for o in variants[1:]]:
        a, ms_a = out[("lanes", nb)]; b, ms_b = out[(other, nb)]
        n = nb * ppb
        assert a[:, :scene.n_bins].sum() == n, "lanes kernel lost photons"
        assert b[:, :scene.n_bins].sum() == n, f"sorted kernel lost photons: {b[:, :scene.n_bins].sum()} vs {n}"
        assert (a[:, :scene.n_bins].sum(1) == ppb).all() and (b[:, :scene.n_bins].sum(1) == ppb).all()
        diff = np.abs(a - b)
        print(f"{label} batches={nb} ppb={ppb}: lanes {ms_a:.2f} ms ({n/ms_a/1e3:.0f} Mph/s)  {other} {ms_b:.2f} ms ({n/ms_b/1e3:.0f} Mph/s)"
              f"  speed-up {ms_a/ms_b:.2f}  bins differing {int((diff[:, :scene.n_bins] > 0).sum())}  photons moved {int(diff[:, :scene.n_bins].sum()) // 2}"
              f"  events {int(a[:, scene.n_bins].sum())} vs {int(b[:, scene.n_bins].sum())}", flush=True)
