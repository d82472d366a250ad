"""Per-class grouping and the report layer (host code over the gathered table).

Replaces nn_inference.py:485-570: the per-class driver loop, the window-3 moving average
rounded to 2 dp (:501-529), the np.histogram summaries (:531-539), the class totals
(:541-558) and ``Results<keyword>_.csv`` (:561-570).  Two defects of the reference are
not reproduced (documented in DESIGN.md): the result lists are never reset between
classes (:463-471 vs :487) and the file name expression ``keywds[k]`` indexes with an
int counter that has reached 9 (:503, :520-521, :570) and raises.
"""
from __future__ import annotations

import csv
from typing import Dict, List, Sequence

import numpy as np

from .schema import (CLASS_KEYWORDS, CLASS_NAMES, CSV_COLUMNS, CSV_KINDS, CSV_SOURCE, FCOL,
                     FLOAT_COLUMNS)


def group_by_class(table, min_contour_area: float = 100.0, num_classes: int = len(CLASS_NAMES)
                   ) -> List[Dict]:
    """One record per class id: instance count (what GetCounts intends, :355-366), number of
    measured rows (contour_area >= cut, :412) and mean / sum of every float column over them."""
    out = []
    cls = table["class_id"]
    ok = (table["valid"] == 1) & (table["contour_area"] >= min_contour_area)
    for k in range(num_classes):
        sel = cls == k
        meas = sel & ok
        rec = dict(class_id=k, class_name=CLASS_NAMES[k] if k < len(CLASS_NAMES) else str(k),
                   keyword=CLASS_KEYWORDS[k] if k < len(CLASS_KEYWORDS) else str(k),
                   count=int(sel.sum()), measured=int(meas.sum()),
                   area_px_sum=int(table["area_px"][sel].sum()))
        f = table.floats[meas]
        for j, c in enumerate(FLOAT_COLUMNS):
            rec["mean_" + c] = float(f[:, j].mean()) if f.shape[0] else 0.0
        out.append(rec)
    return out


def write_classes_csv(path: str, table, min_contour_area: float = 100.0) -> None:
    recs = group_by_class(table, min_contour_area)
    keys = list(recs[0].keys())
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for r in recs:
            w.writerow(r)


_KIND_TYPE = {"f32": np.float32, "f64": np.float64, "py": float}


def moving_average(values: Sequence[float], window_size: int = 3, kind: str = "py") -> List:
    """nn_inference.py:523-527: ``round(sum(window) / window_size, 2)`` of each length-3 window,
    evaluated in the type the reference's list holds (``schema.CSV_KINDS``): np.float32 entries
    sum, divide and round (np.round) in float32, np.float64 entries round with np.round, Python
    floats with the correctly rounded built-in."""
    t = _KIND_TYPE[kind]
    vals = [t(v) for v in values]
    return [round(sum(vals[i:i + window_size]) / window_size, 2)
            for i in range(len(vals) - window_size + 1)]


def smoothed_columns(rows: np.ndarray, window_size: int = 3) -> List[List]:
    """The nine ``MA_*`` lists (:507-529) in CSV column order, entries typed as the reference's."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, len(CSV_COLUMNS))
    return [moving_average(rows[:, j], window_size, CSV_KINDS[j]) for j in range(rows.shape[1])]


def report_class(rows: np.ndarray, window_size: int = 3):
    """rows K x 9 in CSV column order -> (smoothed K-2 x 9 rows, {column: (hist, edges)})."""
    cols = smoothed_columns(rows, window_size)
    sm = np.array([[float(v) for v in c] for c in cols], dtype=np.float64).T.reshape(-1, len(CSV_COLUMNS))
    hists = {name: np.histogram(np.asarray(cols[j])) for j, name in enumerate(CSV_COLUMNS)} \
        if sm.shape[0] else {}
    return sm, hists


def write_shape_descriptor_csv(path: str, rows: np.ndarray, window_size: int = 3) -> None:
    """``ShapeDescriptor.csv`` (:554-559): the smoothed rows through ``csv.writer`` -- the same
    text as the reference's file, float32 entries printed as float32."""
    cols = smoothed_columns(rows, window_size)
    with open(path, "w") as f:
        w = csv.writer(f)
        for row in zip(*cols):
            w.writerow(row)


def write_results_csv(path: str, smoothed_rows: np.ndarray) -> None:
    """``df.to_csv('Results<keyword>_.csv', index=True)`` with the nine named columns (:569-570)."""
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([""] + list(CSV_COLUMNS))
        for i, r in enumerate(np.asarray(smoothed_rows).reshape(-1, len(CSV_COLUMNS))):
            w.writerow([i] + [repr(float(v)) for v in r])
