import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import lscpm_b200; lscpm_b200.install()
from lscpm_b200 import backend, flatten
from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene
arrays = flatten.flatten_scene(standard_scene(tilt=40, elevation=50, azimuth=200, spectrum=am15_like_photon_spectrum()))
pos, direc, wl = backend.emit_rays(arrays, arrays.light, 8, seed=3, device=0)
print(wl, pos[:2], direc[:2])
arrays = flatten.flatten_scene(standard_scene(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()))
pos, direc, wl = backend.emit_rays(arrays, arrays.light, 8, seed=3, device=0)
print(wl)
out = backend.trace_rays(arrays, pos, direc, np.full(8, 500.0), seed=6, device=0, max_steps=8)
print(out["wavelength"][:3], out["event"][:3])
