"""Row schema of the measurement table.

Two tables with identical row order (image-major, then instance order as given):
an int64 table (exact through the all-gather) and a float64 table.  The last nine
float columns are the reference's CSV columns (nn_inference.py:569) computed per
contour exactly as nn_inference.py:434-449 does.
"""

INT_COLUMNS = (
    "image_idx", "inst_idx", "class_id", "valid", "n_contours", "area_px",
    "bbox_x0", "bbox_y0", "bbox_x1", "bbox_y1",
    "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03", "contour_npts",
)
FLOAT_COLUMNS = (
    "score", "cx", "cy", "mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03",
    "equiv_diam_px", "ell_major", "ell_minor", "ell_theta",
    "contour_area", "perimeter",
    "rect_cx", "rect_cy", "rect_w", "rect_h", "rect_angle",
    "Feret", "Aspect_Ratio", "Roundness", "Circularity", "Sphericity",
    "Length", "Width", "CircularED", "Chords",
)
NUM_INT = len(INT_COLUMNS)
NUM_FLOAT = len(FLOAT_COLUMNS)
ICOL = {name: i for i, name in enumerate(INT_COLUMNS)}
FCOL = {name: i for i, name in enumerate(FLOAT_COLUMNS)}

# nn_inference.py:569 -- CSV header of Results<class>_.csv, and the float columns feeding it
CSV_COLUMNS = ("Feret Diameter", "Aspect Ratio", "Roundness", "Circularity", "Sphericity",
               "Length", "Width", "CircularED", "Chords")
CSV_SOURCE = ("Feret", "Aspect_Ratio", "Roundness", "Circularity", "Sphericity",
              "Length", "Width", "CircularED", "Chords")

# Type of the entries the reference appends to its nine lists (nn_inference.py:451-459), in CSV
# column order: order_points returns float32, so everything derived from dA / dB stays np.float32;
# np.sqrt gives np.float64; cv2.contourArea / arcLength arithmetic stays a Python float.  The
# moving average (:523-527) sums, divides and rounds in that type (np.round for the NumPy
# scalars, Python's correctly rounded round() for floats) -- pinned by executing the reference's
# own loop (tests/golden/ref_exec_manifest.json "dtypes").
CSV_KINDS = ("f32", "f32", "f32", "py", "f64", "f32", "f32", "f64", "py")

# nn_inference.py:170 (thing_classes) and :485 (keywds)
CLASS_NAMES = ("Scale bar", "Wall thickness of polyHIPEs", "Pore throats of polyHIPEs",
               "Pores of polyHIPEs")
CLASS_KEYWORDS = ("Scale", "WThick", "PThroat", "Pore")
# nn_inference.py:233 (things_colors)
CLASS_COLORS = ((115, 254, 248), (239, 254, 21), (146, 19, 26), (47, 213, 218))
