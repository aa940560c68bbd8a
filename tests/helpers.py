"""Shared helpers for the parity tests (test code only)."""
from __future__ import annotations

import numpy as np

Z_999 = 3.2905  # two-sided 99.9 % normal quantile


from lscpm_b200.workloads import (AM15_LIKE_W, SPECTRL2_SLICE_NM, ConstantWavelength,  # noqa: E402,F401
                                  am15_like_photon_spectrum, standard_scene)


def bins_outside_band(names, a, b, z=Z_999):
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    assert a.sum() == b.sum(), (a.sum(), b.sum())
    bad = []
    for name, ca, cb in zip(names, a, b):
        if ca == 0 and cb == 0:
            continue
        allowed = z * np.sqrt(max(ca + cb, 4))
        if abs(int(ca) - int(cb)) > allowed:
            bad.append((name, int(ca), int(cb), round(float(allowed), 1)))
    return bad


def assert_counts_agree(names, a, b, z=Z_999, label="", resample=None):
    """Two independent multinomial samples of equal size: every bin's difference inside the 99.9 % band.

    With ~100 bins a 99.9 % band is left by chance in roughly one run out of ten, so a bin that falls outside may be
    re-drawn ONCE with an independent seed (``resample() -> counts``): a real discrepancy fails both draws, a
    fluctuation fails both with probability ~1e-6."""
    bad = bins_outside_band(names, a, b, z)
    if bad and resample is not None:
        again = {name for name, *_ in bins_outside_band(names, resample(), b, z)}
        bad = [entry for entry in bad if entry[0] in again]
    assert not bad, f"{label}: bins outside the 99.9% band: {bad}"


def grouped(names, counts) -> dict:
    """Pool the per-node bins into the families the reference talks about."""
    groups = {}
    for name, c in zip(names, counts):
        parts = name.split("/")
        if parts[0] == "react":
            key = "react"
        elif parts[0] == "absorb":
            key = "absorb/" + ("capillary" if parts[1].startswith("Capillary") else
                               "mixture" if parts[1].startswith("Reaction") else f"plate/{parts[2]}")
        elif parts[0] == "exit" and len(parts) == 3:
            key = "exit/" + ("edge" if parts[2] in ("-x", "+x", "-y", "+y") else parts[2])
        else:
            key = name
        groups[key] = groups.get(key, 0) + int(c)
    return groups


def assert_tallies_agree(names, a, b, label="", resample=None, z=Z_999):
    """Per-bin AND pooled-family agreement of two tally rows with the single re-draw rule of assert_counts_agree."""
    def failing(row):
        bad = {("bin", e[0]): e for e in bins_outside_band(names, row, b, z)}
        g, r = grouped(names, row), grouped(names, b)
        keys = sorted(set(g) | set(r))
        bad.update({("family", e[0]): e for e in bins_outside_band(keys, [g.get(k, 0) for k in keys], [r.get(k, 0) for k in keys], z)})
        return bad

    bad = failing(np.asarray(a, dtype=np.int64))
    if bad and resample is not None:
        again = failing(np.asarray(resample(), dtype=np.int64))
        bad = {k: v for k, v in bad.items() if k in again}
    assert not bad, f"{label}: outside the 99.9% band (also after one independent re-draw): {sorted(bad.values())}"


# ---- the reference's committed per-time-point counts (tests/golden/make_ephemerides.py) ---------------------------
REFERENCE_EPOCH0 = 1577800800   # 2019-12-31T14:00Z


def reference_counts(npz, key):
    """DataFrame of the reference's counts for one result file, indexed by UTC epoch seconds."""
    import pandas as pd

    epoch = REFERENCE_EPOCH0 + 1800 * np.cumsum(npz[key + "|slot"].astype(np.int64))
    columns = {name.split("|")[1]: npz[name].astype(float) for name in npz.files
               if name.startswith(key + "|") and not name.endswith("|slot")}
    return pd.DataFrame(columns, index=epoch)


def frame_epoch(frame):
    """UTC epoch seconds of a tz-aware DatetimeIndex'ed frame"""
    import pandas as pd

    return np.asarray((frame.index.tz_convert("UTC") - pd.Timestamp("1970-01-01", tz="UTC")) // pd.Timedelta(seconds=1))


def binomial_z(k, p, n_ref):
    """z-score of reference counts k_i ~ Binomial(n_ref, p_i) against high-statistics p_i"""
    var = n_ref * np.sum(p * (1 - p))
    return float((k.sum() - n_ref * p.sum()) / np.sqrt(var)) if var > 0 else 0.0
