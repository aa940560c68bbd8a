"""High-statistics GPU vs oracle comparison of pooled tally families (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
import c_oracle
from lscpm_b200 import _native as nat, backend, flatten
from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene
from helpers import grouped
cases = {"585nm": dict(tilt=0, elevation=90, wavelength=585.0)}
_unused = {"config2": dict(tilt=0, elevation=90, spectrum=am15_like_photon_spectrum()), "690nm": dict(tilt=0, elevation=90, wavelength=690.0),
         "585nm": dict(tilt=0, elevation=90, wavelength=585.0)}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
for label, kw in cases.items():
    arrays = flatten.flatten_scene(standard_scene(**kw))
    names = arrays.bin_names()
    gpu = backend.trace_lights(arrays, [arrays.light], n, seed=777, device=0)[0]
    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    b, _ = backend.pack_batches([arrays.light])
    spectra = [arrays.light.spectrum] if arrays.light.spectrum is not None else None
    ref = oracle.trace(b[0], spectra=spectra, photon_count=n, seed=888)
    g, r = grouped(names, gpu[:len(names)]), grouped(names, ref[:len(names)])
    print(label, "events/photon", gpu[len(names)] / n, ref[len(names)] / n, "emissions", gpu[len(names)+1] / n, ref[len(names)+1] / n)
    for nm, a, c in zip(names, gpu, ref):
        if (a or c) and (abs(a - c) / np.sqrt(max(a + c, 1)) > 2.0 or nm.startswith("exit/") or nm in ("kill",)):
            print(f"   BIN {nm:44s} gpu {a/n:.6f} oracle {c/n:.6f}  z = {(a - c) / np.sqrt(max(a + c, 1)):+.2f}")
    for k in sorted(set(g) | set(r)):
        a, c = g.get(k, 0), r.get(k, 0)
        print(f"   {k:32s} gpu {a/n:.5f} oracle {c/n:.5f}  z = {(a - c) / np.sqrt(max(a + c, 1)):+.2f}")
