"""Plate flights on vs off on the GPU at 1e9 photons per arm: every bin, every bin family and the event counters.

The two arms consume the Philox stream differently (a flight draws one free path where the cell march draws several), so
they are statistically independent samples of the same distribution; |z| is the two-sample score per bin.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import lscpm_b200; lscpm_b200.install()
from lscpm_b200 import backend, flatten
from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene
from helpers import grouped

CASES = {
    "config2": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()),
    "config1_585": dict(tilt=0, elevation=90, wavelength=585.0),
    "tilt40_oblique": dict(tilt=40, elevation=35, azimuth=140, spectrum=am15_like_photon_spectrum()),
    "diffuse_tilt30": dict(tilt=30, diffuse=True, spectrum=am15_like_photon_spectrum()),
    "dense46": dict(tilt=0, elevation=60, azimuth=140, wavelength=585.0, capillary_count=46, capillary_pitch=0.01),
    "three_channels_wide": dict(tilt=0, elevation=45, azimuth=90, wavelength=585.0, capillary_count=3, capillary_pitch=0.15),
}
nb, ppb = 16, int(float(sys.argv[1]) / 16) if len(sys.argv) > 1 else 62_500_000
for label, kw in CASES.items():
    arrays = flatten.flatten_scene(standard_scene(**kw))
    names = arrays.bin_names(); n_bins = len(names)
    arms = {}
    for flights in ("0", "1"):
        os.environ["LSCPM_PLATE_FLIGHTS"] = flights
        scene = backend.SceneHandle(arrays, 0)
        plan = backend.Plan(scene, [arrays.light] * nb, list(range(nb)))
        t = plan.new_tallies()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.trace_into(t, ppb, 1000 + int(flights)); e1.record(); torch.cuda.synchronize()
        arms[flights] = (t.cpu().numpy().sum(axis=0), e0.elapsed_time(e1))
    (a, ms_a), (b, ms_b) = arms["0"], arms["1"]
    n = nb * ppb
    assert a[:n_bins].sum() == n and b[:n_bins].sum() == n
    z = (a[:n_bins] - b[:n_bins]) / np.sqrt(np.maximum(a[:n_bins] + b[:n_bins], 1))
    ga, gb = grouped(names, a[:n_bins]), grouped(names, b[:n_bins])
    zg = {k: (ga[k] - gb.get(k, 0)) / np.sqrt(max(ga[k] + gb.get(k, 0), 1)) for k in ga}
    worst = int(np.argmax(np.abs(z)))
    print(f"{label}: {n:.1e} photons/arm  cell march {ms_a:.0f} ms, flights {ms_b:.0f} ms ({ms_a / ms_b:.2f}x)  max|z| bins {np.abs(z).max():.2f} ({names[worst]})"
          f"  families {max(abs(v) for v in zg.values()):.2f}  bins beyond 3.5: {int((np.abs(z) > 3.5).sum())}/{n_bins}"
          f"  events/photon {a[n_bins] / n:.5f} vs {b[n_bins] / n:.5f}  emissions/photon {a[n_bins + 1] / n:.5f} vs {b[n_bins + 1] / n:.5f}", flush=True)
