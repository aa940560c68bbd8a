"""Generates the committed golden fixtures (run here, in the build container; commit the outputs).

The reference's engine (pvtrace) cannot be imported in this image, so these fixtures are produced by the CPU
oracle — the literal Python restatement for per-ray vectors, its C twin for the tallies.  They freeze the
oracle (any drift of oracle/ shows up in the CPU suite) and travel to the GPU box, where /root/reference
does not exist.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import lscpm_b200  # noqa: E402

lscpm_b200.install()
import c_oracle  # noqa: E402
import pvtrace_oracle as po  # noqa: E402
from lscpm_b200 import _native as nat, backend, flatten  # noqa: E402
from lscpm_b200.workloads import am15_like_photon_spectrum, standard_scene  # noqa: E402

TALLY_CASES = {
    "config1_585nm_normal": dict(tilt=0, elevation=90, wavelength=585.0),
    "tilt40_elev50_555nm": dict(tilt=40, elevation=50, wavelength=555.0),
    "grazing_700nm": dict(tilt=0, elevation=10, wavelength=700.0),
    "no_dye_555nm": dict(tilt=40, elevation=50, wavelength=555.0, include_dye=False),
    "diffuse_tilt0_555nm": dict(tilt=0, wavelength=555.0, diffuse=True),
    "diffuse_tilt40_555nm": dict(tilt=40, wavelength=555.0, diffuse=True),
    "config2_am15_spectrum": dict(tilt=0, elevation=90, spectrum="am15"),
}
N_TALLY = 400_000
SEED = 20201


def scene_for(case):
    kw = dict(TALLY_CASES[case])
    if kw.get("spectrum") == "am15":
        kw["spectrum"] = am15_like_photon_spectrum()
    return standard_scene(**kw)


def main():
    tallies = {}
    for case in TALLY_CASES:
        arrays = flatten.flatten_scene(scene_for(case))
        oracle = c_oracle.COracle(nat, nat.PackedScene(arrays))
        batches, _ = backend.pack_batches([arrays.light])
        spectra = [arrays.light.spectrum] if arrays.light.spectrum is not None else None
        row = oracle.trace(batches[0], spectra=spectra, photon_count=N_TALLY, seed=SEED)
        names = arrays.bin_names()
        tallies[case] = {
            "photons": N_TALLY, "seed": SEED, "events": int(row[len(names)]), "emissions": int(row[len(names) + 1]),
            "bins": {n: int(c) for n, c in zip(names, row) if c},
        }
        react = sum(c for n, c in tallies[case]["bins"].items() if n.startswith("react/")) / N_TALLY
        print(f"{case}: reacted {react:.4f} events/photon {row[len(names)] / N_TALLY:.2f}")
    with open(os.path.join(HERE, "oracle_tallies.json"), "w") as fh:
        json.dump({"generator": "oracle/lscpm_oracle.c via tests/golden/make_golden.py", "cases": tallies}, fh, indent=1,
                  sort_keys=True)

    # per-ray vectors from the literal Python oracle (rays harvested along followed photons; FP32-representable)
    out = {}
    for label, kw in (("tilt0", dict(tilt=0, elevation=90, wavelength=585.0)), ("tilt40", dict(tilt=40, elevation=50, wavelength=585.0))):
        scene = standard_scene(**kw)
        oscene = po.OracleScene(scene)
        result = po.run(scene, 150, seed=11, keep_histories=True)
        pos, direc = [], []
        for history in result.histories:
            for step in history[:-1]:
                pos.append(step.position)
                direc.append(step.direction)
        pos = np.array(pos).astype(np.float32).astype(np.float64)
        direc = np.array(direc).astype(np.float32).astype(np.float64)
        direc /= np.linalg.norm(direc, axis=1, keepdims=True)
        rec = dict(hit=[], container=[], adjacent=[], face=[], distance=[], reflectivity=[], reflected=[], refracted=[], normal=[])
        for p, d in zip(pos, direc):
            pr = oscene.probe(p, d)
            if pr is None:
                for k in rec:
                    rec[k].append(-1 if k in ("hit", "container", "adjacent", "face") else (np.zeros(3) if k in ("reflected", "refracted", "normal") else 0.0))
                continue
            rec["hit"].append(pr["hit"]); rec["container"].append(pr["container"])
            rec["adjacent"].append(-1 if pr["adjacent"] is None else pr["adjacent"]); rec["face"].append(pr["face"])
            rec["distance"].append(pr["distance"])
            rec["reflectivity"].append(pr.get("reflectivity", 0.0))
            rec["reflected"].append(pr.get("reflected", np.zeros(3)))
            rec["refracted"].append(np.zeros(3) if pr.get("refracted") is None else pr["refracted"])
            rec["normal"].append(pr.get("normal", np.zeros(3)))
        out[f"{label}_pos"] = pos
        out[f"{label}_dir"] = direc
        for k, v in rec.items():
            out[f"{label}_{k}"] = np.asarray(v)
        print(label, len(pos), "rays")
    np.savez_compressed(os.path.join(HERE, "oracle_probe_vectors.npz"), **out)


if __name__ == "__main__":
    main()
