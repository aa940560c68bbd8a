"""Diagnostic: classify disagreements between lscpm_probe and the oracle's next_hit on rays harvested from oracle
histories, over several scene / light cases."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
import c_oracle
from lscpm_b200 import _native as nat, backend, flatten
from lscpm_b200.workloads import standard_scene, am15_like_photon_spectrum
cases = {
    "oblique": dict(tilt=0, elevation=25, azimuth=100, wavelength=585.0),
    "tilt40": dict(tilt=40, elevation=50, wavelength=555.0),
    "diffuse": dict(tilt=20, wavelength=600.0, diffuse=True),
    "pv_both": dict(tilt=0, elevation=70, azimuth=120, wavelength=585.0, add_bottom_PV=True, add_side_PV=True),
    "pv_backlit": dict(tilt=0, elevation=20, azimuth=0, wavelength=555.0, add_bottom_PV=True, add_side_PV=True),
    "dense46": dict(tilt=0, elevation=60, azimuth=140, wavelength=585.0, capillary_count=46, capillary_pitch=0.01),
    "no_dye": dict(tilt=0, elevation=40, azimuth=60, wavelength=610.0, include_dye=False),
}
n_photons = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
for label, kw in cases.items():
    arrays = flatten.flatten_scene(standard_scene(**kw))
    oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
    b, _ = backend.pack_batches([arrays.light])
    pos, direc, wl = oracle.emit(b[0], n_photons, seed=5)
    hist = oracle.follow(pos, direc, wl, seed=6, max_steps=48)
    P, D = [], []
    for i in range(len(pos)):
        k = min(hist["n_steps"][i], 48)
        keep = np.isin(hist["event"][i, :k - 1], (0, 1, 2, 4))
        P.append(hist["pos"][i, :k - 1][keep]); D.append(hist["dir"][i, :k - 1][keep])
    P = np.concatenate(P).astype(np.float32).astype(np.float64); D = np.concatenate(D).astype(np.float32).astype(np.float64)
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    ref = oracle.probe(P, D)
    gpu = backend.probe_rays(arrays, P, D, device=0)
    names = arrays.node_name
    def kind(i):
        if i < 0: return "none"
        n = names[i]; return "world" if i == 0 else "plate" if n.startswith("LSC") else "pfa" if n.startswith("Cap") else "mix" if n.startswith("Rea") else n
    clear = ref["distance"] > 1e-6
    bad = clear & ~((gpu["hit"] == ref["hit"]) & (gpu["container"] == ref["container"]) & (gpu["adjacent"] == ref["adjacent"]))
    ok = clear & ~bad & (ref["hit"] > 0)
    rel = np.abs(gpu["distance"][ok] - ref["distance"][ok]) / ref["distance"][ok]
    rerr = np.abs(gpu["reflectivity"][ok] - ref["reflectivity"][ok])
    print(f"{label}: rays {len(P)} clear {clear.sum()} interface mismatches {bad.sum()}  distance rel err p50 {np.quantile(rel,0.5):.1e} p99 {np.quantile(rel,0.99):.1e} "
          f"p99.9 {np.quantile(rel,0.999):.1e}  reflectivity abs err p99 {np.quantile(rerr,0.99):.1e}", flush=True)
    cat = collections.Counter()
    for i in np.where(bad)[0]:
        cat[(kind(ref["container"][i]), kind(ref["hit"][i]), kind(ref["adjacent"][i]), "|", kind(gpu["container"][i]), kind(gpu["hit"][i]), kind(gpu["adjacent"][i]))] += 1
    for k, v in cat.most_common(6): print("    ", v, k)
    for i in np.where(bad)[0][:3]:
        print("     e.g.", P[i], D[i], "ref d", ref["distance"][i], "gpu d", gpu["distance"][i])
