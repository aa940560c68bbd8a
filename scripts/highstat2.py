"""Diagnostic: GPU vs oracle at high statistics on scene-family variants (which feature makes a bin differ?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
import c_oracle
from lscpm_b200 import _native as nat, backend, flatten
from test_gpu_edge_cases import _custom_scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
cases = {"bare": dict(n_channels=0, pitch=0.03), "tubes_only": dict(n_channels=16, pitch=0.03, with_inner=False),
         "full": dict(n_channels=16, pitch=0.03), "one": dict(n_channels=1, pitch=0.03)}
for label, kw in cases.items():
    arrays = flatten.flatten_scene(_custom_scene(**kw))
    names = arrays.bin_names()
    gpu = backend.trace_lights(arrays, [arrays.light], n, seed=99, device=0)[0]
    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    b, _ = backend.pack_batches([arrays.light])
    ref = oracle.trace(b[0], photon_count=n, seed=100)
    print(label, "events/photon", gpu[len(names)] / n, ref[len(names)] / n)
    for nm, a, c in zip(names, gpu, ref):
        if (a or c) and (abs(a - c) / np.sqrt(max(a + c, 1)) > 2.5 or (nm.startswith("exit/plate"))):
            print(f"   {nm:30s} gpu {a/n:.6f} oracle {c/n:.6f}  z = {(a - c) / np.sqrt(max(a + c, 1)):+.2f}")
