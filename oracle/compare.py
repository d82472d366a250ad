"""The parity rule of the measurement table (TEST INFRASTRUCTURE; used by tests/ and smoke()).

Integer columns (masks' pixel areas, bboxes, raw moments, contour counts / points, indices):
bit-exact.  Float columns: within ``RTOL`` = 1e-5 relative (the north_star tolerance), column by
column and row by row -- no "fraction of cells" allowance; columns that cancel to ~0 (third-order
central moments of symmetric blobs, mu11, the ellipse angle) get an absolute floor that follows
the rounding noise of their own subtraction."""
import numpy as np

RTOL = 1e-5
INT_COLUMNS = ("image_idx", "inst_idx", "class_id", "valid", "n_contours", "area_px",
               "bbox_x0", "bbox_y0", "bbox_x1", "bbox_y1",
               "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03", "contour_npts")
FLOAT_COLUMNS = ("score", "cx", "cy", "mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03",
                 "equiv_diam_px", "ell_major", "ell_minor", "ell_theta", "contour_area", "perimeter",
                 "rect_cx", "rect_cy", "rect_w", "rect_h", "rect_angle",
                 "Feret", "Aspect_Ratio", "Roundness", "Circularity", "Sphericity",
                 "Length", "Width", "CircularED", "Chords")
IC = {n: i for i, n in enumerate(INT_COLUMNS)}
FC = {n: i for i, n in enumerate(FLOAT_COLUMNS)}


def compare_tables(table, ri, rf, skip_int=()):
    assert table.ints.shape == ri.shape and table.floats.shape == rf.shape
    for name, j in IC.items():
        if name in skip_int:
            continue
        bad = np.flatnonzero(table.ints[:, j] != ri[:, j])
        assert bad.size == 0, f"int column {name}: {bad.size} rows differ, first {bad[:5]}: " \
                              f"{table.ints[bad[:5], j]} vs {ri[bad[:5], j]}"
    exact = 0
    m00 = np.maximum(ri[:, IC["area_px"]].astype(np.float64), 1.0)
    for name, j in FC.items():
        a, b = table.floats[:, j], rf[:, j]
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        ok = ~np.isnan(b)
        a, b = a[ok], b[ok]
        exact += int(np.array_equal(a, b))
        scale = np.abs(b)
        if name in ("mu30", "mu21", "mu12", "mu03"):
            # third-order central moments cancel to ~0 for symmetric blobs: floor at the
            # rounding noise of the subtraction (|m_pq| * eps) relative to m00 * r^3
            scale = np.maximum(scale, m00[ok] ** 2.5 * 1e-6)
        if name in ("mu11", "ell_theta"):
            scale = np.maximum(scale, 1e-6 * (m00[ok] ** 2 if name == "mu11" else 1.0))
        err = np.abs(a - b)
        lim = RTOL * scale + 1e-300
        bad = np.flatnonzero(err > lim)
        assert bad.size == 0, f"float column {name}: {bad.size} rows off, first {bad[:5]}: " \
                              f"{a[bad[:5]]} vs {b[bad[:5]]}"
    return exact
