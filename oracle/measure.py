"""Oracle restatement of the reference's measurement block (TEST INFRASTRUCTURE).

Follows /root/reference/nn_inference.py:
  * ``midpoint``                      :339-340
  * ``get_counts``  (GetCounts)       :355-366
  * ``get_mask_contours`` (GetMask_Contours) :371-459 -- class filter, union paint,
    gray, external contours, left-to-right sort, per-contour descriptor block
  * ``moving_average`` / ``report_class`` :500-570
and adds the per-instance row table of SURVEY.md section 8(b) (the north_star
contract), built from the same OpenCV calls (``cv2.moments``, ``cv2.findContours``,
``cv2.contourArea``, ``cv2.arcLength``, ``cv2.minAreaRect``, ``cv2.boxPoints``).

Side effects of the reference that produce no data (``Image.save``, ``cv2.imwrite``,
``drawContours``/``circle`` on a discarded copy, ``plt.figure``; :398, :402-404,
:417, :422-424) are omitted.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import cv2
import numpy as np
from scipy.spatial import distance as dist

from . import imutils_port as imutils

# nn_inference.py:170 / :485
CLASS_NAMES = ["Scale bar", "Wall thickness of polyHIPEs", "Pore throats of polyHIPEs",
               "Pores of polyHIPEs"]
KEYWORDS = ["Scale", "WThick", "PThroat", "Pore"]
# nn_inference.py:569
CSV_COLUMNS = ['Feret Diameter', 'Aspect Ratio', 'Roundness', 'Circularity', 'Sphericity',
               'Length', 'Width', 'CircularED', 'Chords']

# SURVEY.md section 8(b) row schema (kept independent of the product's copy;
# tests assert the two are identical).
INT_COLUMNS = ["image_idx", "inst_idx", "class_id", "valid", "n_contours", "area_px",
               "bbox_x0", "bbox_y0", "bbox_x1", "bbox_y1",
               "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03", "contour_npts"]
FLOAT_COLUMNS = ["score", "cx", "cy", "mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03",
                 "equiv_diam_px", "ell_major", "ell_minor", "ell_theta",
                 "contour_area", "perimeter",
                 "rect_cx", "rect_cy", "rect_w", "rect_h", "rect_angle",
                 "Feret", "Aspect_Ratio", "Roundness", "Circularity", "Sphericity",
                 "Length", "Width", "CircularED", "Chords"]


def midpoint(ptA, ptB):
    return ((ptA[0] + ptB[0]) * 0.5, (ptA[1] + ptB[1]) * 0.5)


def contour_descriptors(c: np.ndarray, pixels_per_metric: float = 0.85) -> Dict[str, float]:
    """nn_inference.py:414-459 for one contour (no area cut, no list appends)."""
    area = cv2.contourArea(c)
    perimeter = cv2.arcLength(c, True)
    rect = cv2.minAreaRect(c)
    box = cv2.boxPoints(rect)
    box = np.array(box, dtype="int")
    box = imutils.order_points(box)
    (tl, tr, br, bl) = box
    (tltrX, tltrY) = midpoint(tl, tr)
    (blbrX, blbrY) = midpoint(bl, br)
    (tlblX, tlblY) = midpoint(tl, bl)
    (trbrX, trbrY) = midpoint(tr, br)
    dA = dist.euclidean((tltrX, tltrY), (blbrX, blbrY))
    dB = dist.euclidean((tlblX, tlblY), (trbrX, trbrY))
    dimA = dA / pixels_per_metric
    dimB = dB / pixels_per_metric
    dimArea = area / pixels_per_metric
    dimPerimeter = perimeter / pixels_per_metric
    diaFeret = max(dimA, dimB)
    if (dimA and dimB) != 0:
        Aspect_Ratio = max(dimB, dimA) / min(dimA, dimB)
    else:
        Aspect_Ratio = 0
    Length = min(dimA, dimB)
    Width = max(dimA, dimB)
    CircularED = np.sqrt(4 * area / np.pi)
    Chords = cv2.arcLength(c, True)
    Roundness = 1 / (Aspect_Ratio) if Aspect_Ratio != 0 else 0
    with np.errstate(all="ignore"):
        # the reference never reaches perimeter == 0 (area >= 100 cut); IEEE semantics here
        Sphericity = np.float64(2 * np.sqrt(np.pi * dimArea)) / np.float64(dimPerimeter)
        Circularity = 4 * np.pi * (np.float64(dimArea) / np.float64(dimPerimeter) ** 2)
    return dict(contour_area=float(area), perimeter=float(perimeter),
                rect_cx=float(rect[0][0]), rect_cy=float(rect[0][1]),
                rect_w=float(rect[1][0]), rect_h=float(rect[1][1]), rect_angle=float(rect[2]),
                Feret=float(diaFeret), Aspect_Ratio=float(Aspect_Ratio),
                Roundness=float(Roundness), Circularity=float(Circularity),
                Sphericity=float(Sphericity), Length=float(Length), Width=float(Width),
                CircularED=float(CircularED), Chords=float(Chords),
                contour_npts=int(len(c)))


def union_paint(mask_array: np.ndarray, im_shape: Tuple[int, int, int],
                literal: bool = False) -> np.ndarray:
    """nn_inference.py:394-401: OR of the selected masks painted as 255 on a
    zero H x W x 3 uint8 image.  ``literal`` runs the reference's np.where loop
    (O(N*H*W*3) traffic -- used for the timed CPU baseline)."""
    num_instances = mask_array.shape[0]
    output = np.zeros(im_shape, dtype=np.uint8)
    if literal:
        mask_hw_n = np.moveaxis(mask_array, 0, -1)
        for i in range(num_instances):
            output = np.where(mask_hw_n[:, :, i:(i + 1)] == True, 255, output)  # noqa: E712
        return output.astype(np.uint8)
    output[np.any(mask_array, axis=0)] = 255
    return output


def get_mask_contours(im_shape: Tuple[int, int, int], pred_classes: np.ndarray,
                      pred_masks: np.ndarray, classes_of_interest: Sequence[int],
                      min_contour_area: float = 100, pixels_per_metric: float = 0.85,
                      literal_paint: bool = False) -> Optional[np.ndarray]:
    """GetMask_Contours as a pure function.

    Returns K x 9 float64 rows in CSV column order (``CSV_COLUMNS``), one per
    external contour of the class-union image with contourArea >= the cut, in
    left-to-right order; ``None`` when no instance matches (reference prints and
    returns, :383-385).  Raises ValueError when masks exist but are all-false
    (imutils sort_contours on an empty list)."""
    selected_indices = [i for i, cls in enumerate(pred_classes) if cls in classes_of_interest]
    mask_array = pred_masks[selected_indices]
    if mask_array.size == 0:
        return None
    output = union_paint(mask_array, im_shape, literal=literal_paint)
    im_mask = cv2.cvtColor(output, cv2.COLOR_BGR2GRAY)
    cnts = cv2.findContours(im_mask.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    cnts = imutils.grab_contours(cnts)
    (cnts, _) = imutils.sort_contours(cnts)
    rows = []
    for c in cnts:
        if cv2.contourArea(c) < min_contour_area:
            continue
        d = contour_descriptors(c, pixels_per_metric)
        rows.append([d["Feret"], d["Aspect_Ratio"], d["Roundness"], d["Circularity"],
                     d["Sphericity"], d["Length"], d["Width"], d["CircularED"], d["Chords"]])
    return np.array(rows, dtype=np.float64).reshape(-1, 9)


def get_counts(pred_classes: np.ndarray) -> Dict[str, int]:
    """GetCounts (:355-366) literally: ids 1..4 although classes are 0..3, and
    ``PCount`` duplicates ``classes == 3``.  ``intended`` holds the per-class
    histogram the function is meant to produce (ids 0..3)."""
    classes = np.asarray(pred_classes)
    out = dict(TotalCount=int(sum(classes == 1) + sum(classes == 2) + sum(classes == 3)
                              + sum(classes == 4)),
               SCount=int(sum(classes == 1)), WTCount=int(sum(classes == 2)),
               PTCount=int(sum(classes == 3)), PCount=int(sum(classes == 3)))
    out["intended"] = [int((classes == k).sum()) for k in range(len(CLASS_NAMES))]
    return out


# --------------------------------------------------------------------------
# per-instance table (north_star contract)
# --------------------------------------------------------------------------

def _ellipse_from_moments(m00, mu20, mu11, mu02):
    a = mu20 / m00
    b = mu11 / m00
    c = mu02 / m00
    common = math.sqrt(((a - c) * 0.5) ** 2 + b * b)
    lp = (a + c) * 0.5 + common
    lm = (a + c) * 0.5 - common
    major = 4.0 * math.sqrt(max(lp, 0.0))
    minor = 4.0 * math.sqrt(max(lm, 0.0))
    theta = 0.5 * math.atan2(2.0 * b, a - c)
    return major, minor, theta


def frame_moments(mask_u8: np.ndarray, x_off: int = 0, y_off: int = 0) -> Dict[str, float]:
    """cv2.moments(binaryImage=True) of a window placed at (x_off, y_off) in the frame."""
    w = cv2.moments(mask_u8, binaryImage=True)
    r = {k: int(round(w[k])) for k in ("m00", "m10", "m01", "m20", "m11", "m02",
                                       "m30", "m21", "m12", "m03")}
    dx, dy = int(x_off), int(y_off)
    m = {}
    m["m00"] = r["m00"]
    m["m10"] = r["m10"] + dx * r["m00"]
    m["m01"] = r["m01"] + dy * r["m00"]
    m["m20"] = r["m20"] + 2 * dx * r["m10"] + dx * dx * r["m00"]
    m["m11"] = r["m11"] + dx * r["m01"] + dy * r["m10"] + dx * dy * r["m00"]
    m["m02"] = r["m02"] + 2 * dy * r["m01"] + dy * dy * r["m00"]
    m["m30"] = r["m30"] + 3 * dx * r["m20"] + 3 * dx * dx * r["m10"] + dx ** 3 * r["m00"]
    m["m21"] = (r["m21"] + dy * r["m20"] + 2 * dx * r["m11"] + 2 * dx * dy * r["m10"]
                + dx * dx * r["m01"] + dx * dx * dy * r["m00"])
    m["m12"] = (r["m12"] + dx * r["m02"] + 2 * dy * r["m11"] + 2 * dx * dy * r["m01"]
                + dy * dy * r["m10"] + dx * dy * dy * r["m00"])
    m["m03"] = r["m03"] + 3 * dy * r["m02"] + 3 * dy * dy * r["m01"] + dy ** 3 * r["m00"]
    if m["m00"] == 0:
        for k in ("mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03"):
            m[k] = 0.0
        return m
    f = {k: float(v) for k, v in m.items()}
    inv_m00 = 1.0 / f["m00"]
    cx = f["m10"] * inv_m00
    cy = f["m01"] * inv_m00
    mu20 = f["m20"] - f["m10"] * cx
    mu11 = f["m11"] - f["m10"] * cy
    mu02 = f["m02"] - f["m01"] * cy
    m["mu20"], m["mu11"], m["mu02"] = mu20, mu11, mu02
    m["mu30"] = f["m30"] - cx * (3 * mu20 + cx * f["m10"])
    mu11 += mu11
    m["mu21"] = f["m21"] - cx * (mu11 + cx * f["m01"]) - cy * mu20
    m["mu12"] = f["m12"] - cy * (mu11 + cy * f["m10"]) - cx * mu02
    m["mu03"] = f["m03"] - cy * (3 * mu02 + cy * f["m01"])
    return m


def pick_best_contour(cnts_cv_order: Sequence[np.ndarray]) -> Tuple[Optional[np.ndarray], int]:
    """Largest contourArea; ties go to the contour whose start pixel comes first in
    raster order (cv2 returns contours in reverse raster order of their start)."""
    best = None
    best_area = -1.0
    for c in list(cnts_cv_order)[::-1]:
        a = cv2.contourArea(c)
        if a > best_area:
            best, best_area = c, a
    return best, len(cnts_cv_order)


def instance_row(mask_u8: np.ndarray, x_off: int = 0, y_off: int = 0,
                 pixels_per_metric: float = 0.85):
    """One instance mask (window at offset x_off, y_off of the full frame) ->
    (int dict, float dict) of the mask-derived columns in full-frame coordinates."""
    irow = {k: 0 for k in INT_COLUMNS}
    frow = {k: 0.0 for k in FLOAT_COLUMNS}
    irow["bbox_x0"] = irow["bbox_y0"] = irow["bbox_x1"] = irow["bbox_y1"] = -1
    nz = cv2.findNonZero(mask_u8)
    if nz is None:
        return irow, frow
    # Raw moments: cv2.moments on the window (exact integers in double), shifted
    # exactly to full-frame coordinates with Python ints; central moments by
    # OpenCV's completeMomentState formulae (UPSTREAM imgproc/src/moments.cpp) in
    # float64.  tests/test_oracle_measure.py pins this against cv2.moments on the
    # zero-padded full frame.
    m = frame_moments(mask_u8, x_off, y_off)
    irow["valid"] = 1
    irow["area_px"] = int(m["m00"])
    bx, by, bw, bh = cv2.boundingRect(nz)
    irow["bbox_x0"], irow["bbox_y0"] = bx + x_off, by + y_off
    irow["bbox_x1"], irow["bbox_y1"] = bx + x_off + bw - 1, by + y_off + bh - 1
    for k in ("m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03"):
        irow[k] = int(m[k])
    frow["cx"] = m["m10"] / m["m00"]
    frow["cy"] = m["m01"] / m["m00"]
    for k in ("mu20", "mu11", "mu02", "mu30", "mu21", "mu12", "mu03"):
        frow[k] = m[k]
    frow["equiv_diam_px"] = math.sqrt(4.0 * m["m00"] / math.pi)
    frow["ell_major"], frow["ell_minor"], frow["ell_theta"] = _ellipse_from_moments(
        m["m00"], m["mu20"], m["mu11"], m["mu02"])
    cnts, _ = cv2.findContours(mask_u8, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE,
                               offset=(x_off, y_off))
    best, n = pick_best_contour(cnts)
    irow["n_contours"] = n
    d = contour_descriptors(best, pixels_per_metric)
    irow["contour_npts"] = d.pop("contour_npts")
    frow.update(d)
    return irow, frow


def instance_rows(bool_masks: Iterable, classes: np.ndarray, scores: np.ndarray,
                  image_idx: int = 0, inst_idx0: int = 0,
                  pixels_per_metric: float = 0.85) -> Tuple[np.ndarray, np.ndarray]:
    """bool_masks yields either H x W arrays or (window, y_off, x_off) triples."""
    I, F = [], []
    for j, mk in enumerate(bool_masks):
        if isinstance(mk, tuple):
            win, y_off, x_off = mk
        else:
            win, y_off, x_off = mk, 0, 0
        irow, frow = instance_row(np.ascontiguousarray(win, dtype=np.uint8), x_off, y_off,
                                  pixels_per_metric)
        irow["image_idx"] = image_idx
        irow["inst_idx"] = inst_idx0 + j
        irow["class_id"] = int(classes[j])
        frow["score"] = float(scores[j])
        I.append([irow[k] for k in INT_COLUMNS])
        F.append([frow[k] for k in FLOAT_COLUMNS])
    return (np.array(I, dtype=np.int64).reshape(-1, len(INT_COLUMNS)),
            np.array(F, dtype=np.float64).reshape(-1, len(FLOAT_COLUMNS)))


# --------------------------------------------------------------------------
# report layer  (nn_inference.py:500-570)
# --------------------------------------------------------------------------

# Entry types of the nine lists in CSV column order, as the reference's own code produces them
# (observed by executing nn_inference.py:371-459, oracle/ref_exec.py): imutils' order_points
# returns float32, so dA / dB and everything divided from them is np.float32; ``np.sqrt`` yields
# np.float64; the rest stays a Python float.
CSV_ENTRY_TYPES = [np.float32, np.float32, np.float32, float, np.float64,
                   np.float32, np.float32, np.float64, float]


def moving_average(lst: Sequence[float], window_size: int = 3) -> List[float]:
    """:523-527 -- window mean rounded to 2 dp with ``round`` (np.round for NumPy scalars,
    Python's for floats); the arithmetic runs in the type of the list entries."""
    out = []
    i = 0
    while i < (len(lst) - window_size + 1):
        window = lst[i: i + window_size]
        out.append(round(sum(window) / window_size, 2))
        i = i + 1
    return out


def typed_lists(rows: np.ndarray) -> List[list]:
    """K x 9 float64 rows -> the nine lists holding the reference's scalar types."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, len(CSV_COLUMNS))
    return [[t(v) for v in rows[:, j]] for j, t in enumerate(CSV_ENTRY_TYPES)]


def report_class(rows: np.ndarray, window_size: int = 3):
    """rows K x 9 in CSV column order -> (smoothed rows K' x 9, {column: np.histogram})."""
    cols = [moving_average(lst, window_size) for lst in typed_lists(rows)]
    sm = np.array([[float(v) for v in c] for c in cols], dtype=np.float64).T.reshape(-1, len(CSV_COLUMNS))
    hists = {}
    for j, name in enumerate(CSV_COLUMNS):
        if sm.shape[0]:
            hists[name] = np.histogram(np.asarray(cols[j]))
    return sm, hists


def shape_descriptor_text(rows: np.ndarray, window_size: int = 3) -> str:
    """The text of ShapeDescriptor.csv (:554-559) for the rows of one class keyword."""
    import csv
    import io
    cols = [moving_average(lst, window_size) for lst in typed_lists(rows)]
    buf = io.StringIO()
    w = csv.writer(buf)
    for row in zip(*cols):
        w.writerow(row)
    return buf.getvalue().replace("\r\n", "\n")     # as read back in text mode
