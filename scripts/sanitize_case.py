"""Small end-to-end invocation of every kernel for compute-sanitizer (memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lscpm_b200; lscpm_b200.install()
from lscpm_b200 import backend, flatten
from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene
for kw in (dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()), dict(tilt=40, wavelength=555.0, diffuse=True),
           dict(tilt=40, elevation=50, wavelength=555.0, add_bottom_PV=True, add_side_PV=True)):
    arrays = flatten.flatten_scene(standard_scene(**kw))
    rows = backend.trace_lights(arrays, [arrays.light] * 3, 20_000, seed=1, device=0)
    assert np.all(rows[:, :-2].sum(axis=1) == 20_000)
    pos, direc, wl = backend.emit_rays(arrays, arrays.light, 2000, seed=2, device=0)
    out = backend.trace_rays(arrays, pos, direc, wl, seed=3, device=0, max_steps=32)
    probe = backend.probe_rays(arrays, out["pos"][:, 0], out["dir"][:, 0], device=0)
    print(kw.keys(), rows[0, :-2].sum(), int(out["n_steps"].max()), int((probe["hit"] >= 0).sum()))
print("sanitize case ok")
