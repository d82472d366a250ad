"""Pure-Python restatement of OpenCV's external border following (TEST INFRASTRUCTURE).

Small cases only.  Restates, from the published algorithm (Suzuki & Abe 1985 as
implemented by ``cv2.findContours(..., RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)``,
called at nn_inference.py:406), the pieces the CUDA border-trace kernel must
reproduce:

* raster scan for outer-border starts, with the "last marked pixel to the left
  is positive => nested => skip" rule of RETR_EXTERNAL;
* 8-connected border following, first move counter-clockwise on screen;
* CHAIN_APPROX_SIMPLE vertex emission (a point is written when the outgoing
  direction differs from the previous outgoing direction);
* ``contourArea`` (shoelace, exact integers), ``arcLength`` (float32 segment
  lengths summed in float64), ``convexHull`` point set and ``minAreaRect``
  via OpenCV's float32 rotating calipers.

``tests/test_oracle_contours.py`` pins it against the installed cv2.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

# direction code -> (dx, dy); 0 = east, counter-clockwise on a y-up plane,
# i.e. 1 = north-east (x+1, y-1) in image coordinates.
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def find_external_contours(mask: np.ndarray, simple: bool = True):
    """Return list of (points Kx2 int array [x, y], start (x, y)) in RASTER order of
    their start pixel (cv2 returns the reverse of this order)."""
    h, w = mask.shape
    img = np.zeros((h + 2, w + 2), dtype=np.int8)
    img[1:-1, 1:-1] = (mask != 0).astype(np.int8)
    out = []
    for y in range(1, h + 1):
        last_mark = 0  # value of the last marked pixel seen in this row (0 = none)
        x = 1
        while x <= w:
            p = img[y, x]
            if p == 1 and img[y, x - 1] == 0:
                # candidate outer-border start: external iff the last marked pixel
                # to the left is absent or negative ("exited to the right").
                if last_mark <= 0:
                    pts = _follow(img, x, y, simple)
                    out.append((np.array(pts, dtype=np.int32) - 1, (x - 1, y - 1)))
                    p = img[y, x]
            if p != 0 and p != 1:
                last_mark = p
            x += 1
    return out


def _follow(img: np.ndarray, x0: int, y0: int, simple: bool) -> List[Tuple[int, int]]:
    pts: List[Tuple[int, int]] = []
    NBD = 2
    NEG = -126  # (schar)(2 | -128)
    # first search: clockwise from west
    s_end = s = 4
    while True:
        s = (s - 1) & 7
        if img[y0 + DY[s], x0 + DX[s]] != 0 or s == s_end:
            break
    if s == s_end:
        img[y0, x0] = NEG
        pts.append((x0, y0))
        return pts
    i1 = (x0 + DX[s], y0 + DY[s])
    x3, y3 = x0, y0
    prev_s = s ^ 4
    while True:
        s_end = s
        while True:
            s += 1
            x4, y4 = x3 + DX[s & 7], y3 + DY[s & 7]
            if img[y4, x4] != 0:
                break
        s &= 7
        if ((s - 1) & 0xFFFFFFFF) < s_end:
            img[y3, x3] = NEG
        elif img[y3, x3] == 1:
            img[y3, x3] = NBD
        if s != prev_s or not simple:
            pts.append((x3, y3))
            prev_s = s
        if (x4, y4) == (x0, y0) and (x3, y3) == i1:
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return pts


def contour_area(pts: np.ndarray) -> float:
    """cv2.contourArea (unsigned): shoelace over the vertex list."""
    p = np.asarray(pts, dtype=np.int64).reshape(-1, 2)
    if len(p) == 0:
        return 0.0
    q = np.roll(p, 1, axis=0)
    a = int(np.sum(q[:, 0] * p[:, 1] - q[:, 1] * p[:, 0]))
    return abs(a) * 0.5


def arc_length(pts: np.ndarray) -> float:
    """cv2.arcLength(c, True): float32 sqrt per segment, double accumulation."""
    p = np.asarray(pts, dtype=np.float32).reshape(-1, 2)
    if len(p) <= 1:
        return 0.0
    q = np.roll(p, 1, axis=0)
    d = p - q
    seg = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)
    per = 0.0
    for v in seg:
        per += float(v)
    return per


def convex_hull(pts: np.ndarray) -> np.ndarray:
    """Strict convex hull (no collinear points), Andrew monotone chain.

    Orientation equals cv2.convexHull(clockwise=False): clockwise ON SCREEN
    (x right, y down).  Start point: see ``hull_like_cv``.
    """
    p = np.unique(np.asarray(pts, dtype=np.int64).reshape(-1, 2), axis=0)
    p = p[np.lexsort((p[:, 1], p[:, 0]))]
    if len(p) <= 2:
        return p

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower = []
    for q in p:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], q) <= 0:
            lower.pop()
        lower.append(tuple(q))
    upper = []
    for q in p[::-1]:
        while len(upper) >= 2 and cross(upper[-2], upper[-1], q) <= 0:
            upper.pop()
        upper.append(tuple(q))
    return np.array(lower[:-1] + upper[:-1], dtype=np.int64)


def rotating_calipers_f32(hull: np.ndarray):
    """OpenCV rotcalipers.cpp::rotatingCalipers(CALIPERS_MINAREARECT) in float32.

    ``hull`` is n x 2 (n >= 3) in the order cv2.convexHull produced it.
    Returns out[6] float32: corner (px, py), vec1 (x, y), vec2 (x, y).
    """
    f32 = np.float32
    pts = np.asarray(hull, dtype=f32)
    n = len(pts)
    vect = np.zeros((n, 2), dtype=f32)
    inv_len = np.zeros(n, dtype=f32)
    left = bottom = right = top = 0
    left_x = right_x = pts[0, 0]
    top_y = bottom_y = pts[0, 1]
    pt0 = pts[0]
    for i in range(n):
        if pt0[0] < left_x:
            left_x, left = pt0[0], i
        if pt0[0] > right_x:
            right_x, right = pt0[0], i
        if pt0[1] > top_y:
            top_y, top = pt0[1], i
        if pt0[1] < bottom_y:
            bottom_y, bottom = pt0[1], i
        pt = pts[(i + 1) % n]
        dx = float(f32(pt[0] - pt0[0]))
        dy = float(f32(pt[1] - pt0[1]))
        vect[i, 0] = f32(dx)
        vect[i, 1] = f32(dy)
        inv_len[i] = f32(1.0 / math.sqrt(dx * dx + dy * dy))
        pt0 = pt
    orientation = f32(0)
    ax, ay = float(vect[n - 1, 0]), float(vect[n - 1, 1])
    for i in range(n):
        bx, by = float(vect[i, 0]), float(vect[i, 1])
        convexity = ax * by - ay * bx
        if convexity != 0:
            orientation = f32(1) if convexity > 0 else f32(-1)
            break
        ax, ay = bx, by
    assert orientation != 0
    base_a = orientation
    base_b = f32(0)
    seq = [bottom, right, top, left]
    minarea = f32(np.finfo(np.float32).max)
    buf = None
    for _k in range(n):
        # choose the caliper side with the minimum angle to its polygon edge by the
        # sign of an (exact) cross product: rot_vect[i] = edge i rotated into side 0's frame
        rv = [
            (vect[seq[0], 0], vect[seq[0], 1]),
            (vect[seq[1], 1], -vect[seq[1], 0]),      # rotate90CW
            (-vect[seq[2], 0], -vect[seq[2], 1]),     # rotate180
            (-vect[seq[3], 1], vect[seq[3], 0]),      # rotate90CCW
        ]
        main = 0
        for i in range(1, 4):
            # firstVecIsRight(rv[i], rv[main]): rotate90CW(v1) . v2 < 0
            t0, t1 = rv[i][1], -rv[i][0]
            if f32(f32(t0 * rv[main][0]) + f32(t1 * rv[main][1])) < 0:
                main = i
        pindex = seq[main]
        lead_x = f32(vect[pindex, 0] * inv_len[pindex])
        lead_y = f32(vect[pindex, 1] * inv_len[pindex])
        if main == 0:
            base_a, base_b = lead_x, lead_y
        elif main == 1:
            base_a, base_b = lead_y, f32(-lead_x)
        elif main == 2:
            base_a, base_b = f32(-lead_x), f32(-lead_y)
        else:
            base_a, base_b = f32(-lead_y), lead_x
        seq[main] += 1
        if seq[main] == n:
            seq[main] = 0
        dx = f32(pts[seq[1], 0] - pts[seq[3], 0])
        dy = f32(pts[seq[1], 1] - pts[seq[3], 1])
        width = f32(f32(dx * base_a) + f32(dy * base_b))
        dx = f32(pts[seq[2], 0] - pts[seq[0], 0])
        dy = f32(pts[seq[2], 1] - pts[seq[0], 1])
        height = f32(f32(-dx * base_b) + f32(dy * base_a))
        area = f32(width * height)
        if area <= minarea:  # keeps the LAST minimum
            minarea = area
            buf = (seq[3], base_a, width, base_b, height, seq[0], area)
    l_idx, A1, width, B1, height, b_idx, _ = buf
    A2 = f32(-B1)
    B2 = A1
    C1 = f32(f32(A1 * pts[l_idx, 0]) + f32(pts[l_idx, 1] * B1))
    C2 = f32(f32(A2 * pts[b_idx, 0]) + f32(pts[b_idx, 1] * B2))
    idet = f32(f32(1) / f32(f32(A1 * B2) - f32(A2 * B1)))
    px = f32(f32(f32(C1 * B2) - f32(C2 * B1)) * idet)
    py = f32(f32(f32(A1 * C2) - f32(A2 * C1)) * idet)
    out = np.array([px, py, f32(A1 * width), f32(B1 * width), f32(A2 * height), f32(B2 * height)],
                   dtype=f32)
    return out


def hull_like_cv(contour_pts: np.ndarray) -> np.ndarray:
    """Hull of a traced external contour in the order cv2.convexHull(c, clockwise=False)
    returns it for a simple contour: clockwise on screen, cyclically shifted so
    that the hull indices into the contour descend, i.e. the contour's start
    pixel (raster-first pixel of the component, always a hull vertex) comes LAST.
    (UPSTREAM convhull.cpp: "try to make the convex hull indices form an
    ascending or descending sequence by the cyclic shift of the output".)"""
    pts = np.asarray(contour_pts, dtype=np.int64).reshape(-1, 2)
    h = convex_hull(pts)
    if len(h) <= 2:
        return h[::-1]       # natural order starts at the (x, y)-largest point
    # raster-first point of the contour = min y, then min x
    c0 = pts[np.lexsort((pts[:, 0], pts[:, 1]))][0]
    k = int(np.where((h == c0).all(axis=1))[0][0])
    return np.roll(h, -(k + 1), axis=0)


def min_area_rect_cv(hull: np.ndarray):
    """cv2.minAreaRect restated (UPSTREAM rotcalipers.cpp::minAreaRect), float32 pipeline.

    Returns ((cx, cy), (w, h), angle_deg) as Python floats holding float32 values,
    with OpenCV >= 4.5.1's angle convention (angle in [-90, 0), sides swapped to match).
    """
    f32 = np.float32
    hull = np.asarray(hull).reshape(-1, 2)
    n = len(hull)
    if n > 2:
        out = rotating_calipers_f32(hull)
        cx = f32(out[0] + f32(f32(out[2] + out[4]) * f32(0.5)))
        cy = f32(out[1] + f32(f32(out[3] + out[5]) * f32(0.5)))
        w = f32(math.sqrt(float(out[2]) * float(out[2]) + float(out[3]) * float(out[3])))
        h = f32(math.sqrt(float(out[4]) * float(out[4]) + float(out[5]) * float(out[5])))
        ang = math.atan2(float(out[3]), float(out[2])) * 180.0 / math.pi
        return ((float(cx), float(cy)),) + _normalise(w, h, ang)
    if n == 2:
        p = hull.astype(f32)
        cx = f32(f32(p[0, 0] + p[1, 0]) * f32(0.5))
        cy = f32(f32(p[0, 1] + p[1, 1]) * f32(0.5))
        dx = float(f32(p[1, 0] - p[0, 0]))
        dy = float(f32(p[1, 1] - p[0, 1]))
        w = f32(math.sqrt(dx * dx + dy * dy))
        h = f32(0)
        ang = math.atan2(dy, dx) * 180.0 / math.pi
        return ((float(cx), float(cy)),) + _normalise(w, h, ang)
    if n == 1:
        return ((float(hull[0, 0]), float(hull[0, 1])),) + _normalise(f32(0), f32(0), 0.0)
    return ((0.0, 0.0),) + _normalise(f32(0), f32(0), 0.0)


def _normalise(w, h, ang):
    """OpenCV >= 4.5.1 convention, pinned by probe: angle (double) moved into [-90, 0)
    in 90-degree steps, swapping the sides at every step, then cast to float32."""
    while ang >= 0.0:
        ang -= 90.0
        w, h = h, w
    while ang < -90.0:
        ang += 90.0
        w, h = h, w
    return ((float(w), float(h)), float(np.float32(ang)))


def box_points_cv(rect) -> np.ndarray:
    """cv2.boxPoints restated (UPSTREAM RotatedRect::points), 4 x 2 float32."""
    f32 = np.float32
    (cx, cy), (w, h), angle = rect
    cx, cy, w, h = f32(cx), f32(cy), f32(w), f32(h)
    _angle = float(f32(angle)) * math.pi / 180.0
    b = f32(f32(math.cos(_angle)) * f32(0.5))
    a = f32(f32(math.sin(_angle)) * f32(0.5))
    pt = np.zeros((4, 2), dtype=f32)
    pt[0, 0] = f32(f32(cx - f32(a * h)) - f32(b * w))
    pt[0, 1] = f32(f32(cy + f32(b * h)) - f32(a * w))
    pt[1, 0] = f32(f32(cx + f32(a * h)) - f32(b * w))
    pt[1, 1] = f32(f32(cy - f32(b * h)) - f32(a * w))
    pt[2, 0] = f32(f32(f32(2) * cx) - pt[0, 0])
    pt[2, 1] = f32(f32(f32(2) * cy) - pt[0, 1])
    pt[3, 0] = f32(f32(f32(2) * cx) - pt[1, 0])
    pt[3, 1] = f32(f32(f32(2) * cy) - pt[1, 1])
    return pt
