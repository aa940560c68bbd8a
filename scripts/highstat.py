"""High-statistics GPU-vs-oracle sweep over light / scene cases: flags every bin with |z| > 3.5 (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
import c_oracle
from lscpm_b200 import _native as nat, backend, flatten
from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene
from helpers import grouped
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
cases = {
    "tilt40_elev50_555": dict(tilt=40, elevation=50, wavelength=555.0),
    "grazing_700": dict(tilt=0, elevation=10, wavelength=700.0),
    "no_dye_555": dict(tilt=40, elevation=50, wavelength=555.0, include_dye=False),
    "diffuse_tilt0": dict(tilt=0, wavelength=555.0, diffuse=True),
    "diffuse_tilt40": dict(tilt=40, wavelength=555.0, diffuse=True),
    "config2": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()),
    "oblique_az100_610": dict(tilt=-10, elevation=25, azimuth=100, wavelength=610.0),
    "pv_both": dict(tilt=40, elevation=50, wavelength=555.0, add_bottom_PV=True, add_side_PV=True),
    "pv_side_585": dict(tilt=0, elevation=90, wavelength=585.0, add_side_PV=True),
    "dense46": dict(tilt=0, elevation=60, azimuth=140, wavelength=585.0, capillary_count=46, capillary_pitch=0.01),
    "config1_585": dict(tilt=0, elevation=90, wavelength=585.0),
    "tilt_neg10_az213_450": dict(tilt=-10, elevation=35, azimuth=213.5, wavelength=450.0),
    "tilt80_elev15_650": dict(tilt=80, elevation=15, azimuth=170, wavelength=650.0),
    "pv_backlit": dict(tilt=40, elevation=20, azimuth=0, wavelength=555.0, add_bottom_PV=True, add_side_PV=True),
    "pv_bottom_diffuse": dict(tilt=20, wavelength=600.0, diffuse=True),
    "pv_bottom_am15": dict(tilt=0, elevation=70, azimuth=120, spectrum=am15_like_photon_spectrum(), add_bottom_PV=True),
    "no_dye_pv_side": dict(tilt=10, elevation=60, wavelength=610.0, include_dye=False, add_side_PV=True),
    "three_channels_wide": dict(tilt=0, elevation=45, azimuth=90, wavelength=585.0, capillary_count=3, capillary_pitch=0.15),
    "two_channels_wide": dict(tilt=0, elevation=45, azimuth=90, wavelength=585.0, capillary_count=2, capillary_pitch=0.2),
    "three_channels_normal": dict(tilt=0, elevation=90, wavelength=585.0, capillary_count=3, capillary_pitch=0.15),
    "backlit_no_pv": dict(tilt=60, elevation=10, azimuth=20, wavelength=555.0),
}
only = sys.argv[2:] or list(cases)
for label in only:
    kw = cases[label]
    arrays = flatten.flatten_scene(standard_scene(**kw))
    names = arrays.bin_names(); nb = len(names)
    t0 = time.time()
    gpu = backend.trace_lights(arrays, [arrays.light], n, seed=int(os.environ.get('HS_SEED', '4242')), device=0)[0]
    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    b, _ = backend.pack_batches([arrays.light])
    spectra = [arrays.light.spectrum] if arrays.light.spectrum is not None else None
    ref = oracle.trace(b[0], spectra=spectra, photon_count=n, seed=int(os.environ.get('HS_SEED', '4242')) + 7)
    z = (gpu[:nb] - ref[:nb]) / np.sqrt(np.maximum(gpu[:nb] + ref[:nb], 1))
    g, r = grouped(names, gpu[:nb]), grouped(names, ref[:nb])
    zg = {k: (g.get(k, 0) - r.get(k, 0)) / np.sqrt(max(g.get(k, 0) + r.get(k, 0), 1)) for k in set(g) | set(r)}
    flagged = [(names[i], int(gpu[i]), int(ref[i]), round(float(z[i]), 2)) for i in np.where(np.abs(z) > 3.5)[0]]
    flagged += [("family:" + k, g.get(k, 0), r.get(k, 0), round(float(v), 2)) for k, v in zg.items() if abs(v) > 3.5]
    print(f"{label}: max|z| bins {np.abs(z).max():.2f} families {max(abs(v) for v in zg.values()):.2f} "
          f"events/photon gpu {gpu[nb]/n:.5f} oracle {ref[nb]/n:.5f} ({time.time()-t0:.0f}s) flagged: {flagged}", flush=True)
