"""Scratch timing of the trace kernel (not the bench contract): photons/s for a few batch sizes."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import lscpm_b200
lscpm_b200.install()
from lscpm_b200 import backend, flatten
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import standard_scene, am15_like_photon_spectrum

for label, kw in (("585nm", dict(wavelength=585.0)), ("am15", dict(spectrum=am15_like_photon_spectrum()))):
    arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, **kw))
    scene = backend.scene_handle(arrays, 0)
    for nb, ppb in ((1, 1_000_000), (1, 16_000_000), (64, 1_000_000), (1, 256_000_000)):
        plan = backend.Plan(scene, [arrays.light] * nb, list(range(nb)))
        tallies = plan.new_tallies()
        plan.trace_into(tallies, ppb, 1); torch.cuda.synchronize()
        tallies.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.trace_into(tallies, ppb, 2); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = tallies.cpu().numpy()
        n = nb * ppb
        ev = t[:, scene.n_bins].sum()
        print(f"{label} batches={nb} ppb={ppb}: {ms:.2f} ms  {n/ms/1e3:.1f} Mphotons/s  {ev/ms/1e3:.1f} Mevents/s  events/photon {ev/n:.2f} react {t[:, :scene.n_bins][:, [i for i,b in enumerate(scene.bin_names) if b.startswith('react/')]].sum()/n:.4f}", flush=True)
