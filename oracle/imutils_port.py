"""Oracle restatement of the three imutils helpers the reference calls (TEST INFRASTRUCTURE).

imutils 0.5.4 is not installed in this image and not vendored by the reference
(imports at nn_inference.py:18-20; call sites :407, :408, :421).  Restated from
the published source (imutils/convenience.py, contours.py, perspective.py).
"""
from __future__ import annotations

import cv2
import numpy as np


def grab_contours(cnts):
    """imutils.convenience.grab_contours: pick the contour list out of the
    findContours return tuple (2-tuple on OpenCV 2.4/4.x, 3-tuple on 3.x)."""
    if len(cnts) == 2:
        cnts = cnts[0]
    elif len(cnts) == 3:
        cnts = cnts[1]
    else:
        raise Exception(("Contours tuple must have length 2 or 3, "
                         "otherwise OpenCV changed their cv2.findContours return "
                         "signature yet again. Refer to OpenCV's documentation "
                         "in that case"))
    return cnts


def sort_contours(cnts, method="left-to-right"):
    """imutils.contours.sort_contours: stable sort on boundingRect x (or y).
    Raises ValueError on an empty list, exactly as zip(*[]) unpacking does upstream."""
    reverse = False
    i = 0
    if method == "right-to-left" or method == "bottom-to-top":
        reverse = True
    if method == "top-to-bottom" or method == "bottom-to-top":
        i = 1
    boundingBoxes = [cv2.boundingRect(c) for c in cnts]
    (cnts, boundingBoxes) = zip(*sorted(zip(cnts, boundingBoxes),
                                        key=lambda b: b[1][i], reverse=reverse))
    return (cnts, boundingBoxes)


def order_points(pts):
    """imutils.perspective.order_points -> float32 [tl, tr, br, bl]."""
    xSorted = pts[np.argsort(pts[:, 0]), :]
    leftMost = xSorted[:2, :]
    rightMost = xSorted[2:, :]
    leftMost = leftMost[np.argsort(leftMost[:, 1]), :]
    (tl, bl) = leftMost
    # scipy.spatial.distance.cdist(tl[np.newaxis], rightMost, "euclidean")[0]
    d = rightMost.astype(np.float64) - tl.astype(np.float64)[None, :]
    D = np.sqrt((d * d).sum(axis=1))
    (br, tr) = rightMost[np.argsort(D)[::-1], :]
    return np.array([tl, tr, br, bl], dtype="float32")
