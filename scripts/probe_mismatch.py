"""Per-ray parity (check #1) diagnostics: lscpm_probe against the oracle's next_hit on rays harvested from oracle histories.

Prints, per scene / light case: interface mismatches by category, the distribution of the relative distance error, and a
characterisation of its tail — for the worst rays the distance t, |dt|, the cosine of incidence on the surface that is hit
and the kind of surface — plus the largest value over ALL rays of the normalised error

    |dt| / (1e-5 * t  +  K * ulp32(0.25 m) / |cos i|)

i.e. the north-star tolerance plus K units of FP32 resolution of a plate coordinate, measured along the surface normal
(a hit point that is displaced by delta along the normal moves the distance by delta / |cos i|: grazing hits are
ill-conditioned in any 32-bit arithmetic).  tests/test_gpu_probe_parity.py asserts that bound on every ray.

Usage: python scripts/probe_mismatch.py [photons] ; LSCPM_LIBRARY=variants/precise.so selects the IEEE-division build."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lscpm_b200; lscpm_b200.install()
import c_oracle
from lscpm_b200 import _native as nat, backend, flatten
from lscpm_b200.workloads import standard_scene

ULP = float(np.spacing(np.float32(0.25)))   # 2.98e-8 m: FP32 resolution of a coordinate on the 0.47 m plate
K = 4.0

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
print(f"library: {os.environ.get('LSCPM_LIBRARY', 'csrc/liblscpm.so')}   ulp32(0.25 m) = {ULP:.3e}  K = {K}")
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
    t = ref["distance"][ok]; dt = np.abs(gpu["distance"][ok] - t)
    cosi = np.abs(np.einsum("ij,ij->i", ref["normal"][ok], D[ok]))
    rel = dt / t
    bound = 1e-5 * t + K * ULP / np.maximum(cosi, 1e-12)
    norm_err = dt / bound
    rerr = np.abs(gpu["reflectivity"][ok] - ref["reflectivity"][ok])
    print(f"{label}: rays {len(P)} clear {clear.sum()} interface mismatches {bad.sum()}  distance rel err p50 {np.quantile(rel,0.5):.1e} "
          f"p99 {np.quantile(rel,0.99):.1e} p99.9 {np.quantile(rel,0.999):.1e} max {rel.max():.1e}  over 1e-5: {np.mean(rel > 1e-5):.2e} of rays  "
          f"normalised error max {norm_err.max():.3f} p99.99 {np.quantile(norm_err, 0.9999):.3f}  reflectivity abs err p99 {np.quantile(rerr,0.99):.1e}", flush=True)
    tail = np.argsort(rel)[-max(len(rel) // 1000, 5):]
    hitkind = np.array([kind(i) for i in ref["hit"][ok][tail]])
    print(f"    worst 0.1%: median t {np.median(t[tail]):.2e} m, median |dt| {np.median(dt[tail]):.2e} m, median |cos i| {np.median(cosi[tail]):.2e} "
          f"(all rays: {np.median(cosi):.2e});  surfaces {dict(collections.Counter(hitkind))};  |dt|*|cos i| max {np.max(dt[tail] * cosi[tail]):.2e} m")
    for i in tail[-3:]:
        print(f"      t {t[i]:.4e}  dt {dt[i]:.2e}  rel {rel[i]:.2e}  cos i {cosi[i]:.2e}  hit {kind(ref['hit'][ok][i])}  normalised {norm_err[i]:.3f}")
    cat = collections.Counter()
    for i in np.where(bad)[0]:
        cat[(kind(ref["container"][i]), kind(ref["hit"][i]), kind(ref["adjacent"][i]), "|", kind(gpu["container"][i]), kind(gpu["hit"][i]), kind(gpu["adjacent"][i]))] += 1
    for k, v in cat.most_common(6): print("    ", v, k)
