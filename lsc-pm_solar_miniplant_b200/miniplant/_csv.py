"""``DataFrame.to_csv`` for the drivers' result tables, byte for byte, without pandas' per-cell formatting.

The reference writes its results with ``to_csv`` (``miniplant/full_simulation.py:115-125``,
``angle_optimization_simulation.py:93-108``, ``solar_data.py:143-155``): a tz-aware DatetimeIndex plus float columns.
pandas formats both cell by cell (0.13 s per year-long table, more than the whole GPU trace); the same text is produced
here column-wise.  ``tests/test_host_logic.py`` compares the two outputs byte for byte.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _index_strings(index) -> list:
    if not isinstance(index, pd.DatetimeIndex):
        return [str(v) for v in index]
    if index.tz is None:
        local = index
        suffix = None
    else:
        local = index.tz_localize(None)
        offset_min = ((local.values - index.tz_convert("UTC").tz_localize(None).values) // np.timedelta64(1, "m")).astype(np.int64)
        sign = np.where(offset_min < 0, "-", "+")
        mag = np.abs(offset_min)
        suffix = np.char.add(np.char.add(sign, np.char.zfill((mag // 60).astype(str), 2)),
                             np.char.add(":", np.char.zfill((mag % 60).astype(str), 2)))
    values = local.values.astype("datetime64[ns]")
    whole = bool(np.all(values == values.astype("datetime64[s]")))
    text = np.datetime_as_string(values, unit="s" if whole else "ns")
    text = np.char.replace(text, "T", " ")
    if not whole:   # pandas prints microseconds / nanoseconds only as far as needed; leave those tables to pandas
        raise ValueError("sub-second timestamps")
    if suffix is not None:
        text = np.char.add(text, suffix)
    return text.tolist()


def _column_strings(values) -> list:
    arr = np.asarray(values)
    if arr.dtype.kind == "f":
        out = [repr(v) for v in arr.tolist()]
        if np.isnan(arr).any():
            out = ["" if s == "nan" else s for s in out]
        return out
    if arr.dtype.kind in "iu":
        return [str(v) for v in arr.tolist()]
    raise ValueError(f"unsupported column dtype {arr.dtype}")


def write_csv(frame: pd.DataFrame, path, columns) -> None:
    """``frame.to_csv(path, columns=columns)`` (falls back to pandas for anything but float/int columns)."""
    try:
        cols = [_column_strings(frame[c].to_numpy()) for c in columns]
        idx = _index_strings(frame.index)
    except ValueError:
        frame.to_csv(path, columns=columns)
        return
    header = ",".join([frame.index.name or ""] + list(columns))
    lines = [header]
    lines.extend(",".join(row) for row in zip(idx, *cols))
    with open(path, "w", newline="") as fh:
        fh.write("\n".join(lines))
        fh.write("\n")
