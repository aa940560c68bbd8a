"""BASELINE configs[3] as one launch: every half-hour point of the Eindhoven 40 deg year (direct + diffuse batches, own
SPECTRL2 spectra), N photons per batch; device-timed.  Usage: python scripts/config4_timing.py [photons_per_batch=100000]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import lscpm_b200; lscpm_b200.install()
from lscpm_b200 import backend, flatten, lights
from miniplant.locations import EINDHOVEN
from miniplant.scene_creator import create_direct_scene
from miniplant.solar_data import solar_data_for_place_and_time
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
t0 = time.time(); data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False); t_solar = time.time() - t0
t0 = time.time()
batch = [lights.direct_light(40, r["apparent_elevation"], r["azimuth"], r["direct_spectrum"]) for _, r in data.iterrows()]
batch += [lights.diffuse_light(40, r["diffuse_spectrum"]) for _, r in data.iterrows()]
t_lights = time.time() - t0
arrays = flatten.flatten_scene(create_direct_scene(tilt_angle=40, include_dye=True), with_light=False)
scene = backend.scene_handle(arrays, 0)
t0 = time.time(); plan = backend.Plan(scene, batch, list(range(len(batch)))); t_plan = time.time() - t0
tallies = plan.new_tallies()
plan.trace_into(tallies, n, 1); torch.cuda.synchronize()
ms = []
for k in range(3):
    tallies.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.trace_into(tallies, n, 2 + k); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
total = len(batch) * n
print(f"config 4 (Eindhoven 40 deg year): {len(batch)} batches x {n} photons = {total/1e9:.2f}e9 photons per launch: "
      f"{min(ms):.1f} ms = {total/min(ms)/1e6:.2f}e9 photons/s | host: solar stage {t_solar:.2f} s, light descriptors {t_lights:.2f} s, plan {t_plan:.2f} s")
n = 120
tallies.zero_(); plan.trace_into(tallies, n, 9); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tallies.zero_(); e0.record(); plan.trace_into(tallies, n, 10); e1.record(); torch.cuda.synchronize()
print(f"the reference's own setting, {n} photons per point ({len(batch) * n / 1e6:.2f}e6 photons): {e0.elapsed_time(e1):.2f} ms")
