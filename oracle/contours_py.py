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


def _sklansky(P, start: int, end: int, nsign: int, sign2: int) -> List[int]:
    """UPSTREAM convhull.cpp::Sklansky_ over the sorted points P (one quarter of the hull, walked
    from ``start`` towards ``end``); returns the stack of indices into P."""
    def sgn(v):
        return (v > 0) - (v < 0)
    incr = 1 if end > start else -1
    pprev, pcur, pnext = start, start + incr, start + 2 * incr
    if start == end or P[start] == P[end]:
        return [start]
    stack = [pprev, pcur, pnext]
    end += incr
    while pnext != end:
        cury, nexty = P[pcur][1], P[pnext][1]
        by = nexty - cury
        if sgn(by) != nsign:
            ax = P[pcur][0] - P[pprev][0]
            bx = P[pnext][0] - P[pcur][0]
            ay = cury - P[pprev][1]
            convexity = ay * bx - ax * by
            if sgn(convexity) == sign2 and (ax != 0 or ay != 0):
                pprev, pcur = pcur, pnext
                pnext += incr
                stack.append(pnext)
            elif pprev == start:
                pcur = pnext
                stack[1] = pcur
                pnext += incr
                stack[2] = pnext
            else:
                stack[-2] = pnext
                pcur = pprev
                pprev = stack[-4]
                stack.pop()
        else:
            pnext += incr
            stack[-1] = pnext
    stack.pop()
    return stack


def convex_hull_cv(contour_pts: np.ndarray) -> np.ndarray:
    """cv2.convexHull(points, clockwise=False, returnPoints=True) restated IN FULL for integer
    points -- what cv2.minAreaRect (called at nn_inference.py:417) hands to its rotating calipers,
    including the ORDER of the output, which decides ties between rectangles of equal float32 area.

    OpenCV is a dependency of the reference, not part of /root/reference (opencv-python 4.13 in
    this image); restated from its published source (imgproc/src/convhull.cpp): points sorted by
    (x, y, position), four Sklansky walks (top-left, top-right, bottom-left, bottom-right quarter),
    then "try to make the convex hull indices form an ascending or descending sequence by the
    cyclic shift of the output" -- a shift that is only made when the contour indices of the hull
    vertices are a rotation of a monotone sequence.  For a simple contour they always are and the
    result is ``hull_like_cv``'s order; a contour that visits a hull vertex twice (one-pixel spurs)
    can fail the test and keeps the natural order, which starts at the (x, y)-largest point.
    tests/test_oracle.py pins this against cv2.convexHull and cv2.minAreaRect on thousands of
    speckle contours, bit for bit."""
    pts = [(int(x), int(y)) for x, y in np.asarray(contour_pts).reshape(-1, 2)]
    total = len(pts)
    if total == 0:
        return np.zeros((0, 2), dtype=np.int64)
    order = sorted(range(total), key=lambda i: (pts[i][0], pts[i][1], i))
    P = [pts[i] for i in order]
    miny = maxy = 0
    for i in range(1, total):
        if P[miny][1] > P[i][1]:
            miny = i
        if P[maxy][1] < P[i][1]:
            maxy = i
    if P[0] == P[total - 1]:
        return np.array([pts[order[0]]], dtype=np.int64)
    hull: List[int] = []
    # (clockwise=False: the two upper stacks, then the two lower ones, swap places)
    tl = _sklansky(P, total - 1, maxy, -1, -1)
    tr = _sklansky(P, 0, maxy, -1, 1)
    hull += [order[k] for k in tl[:-1]]
    hull += [order[tr[i]] for i in range(len(tr) - 1, 0, -1)]
    stop_idx = tr[1] if len(tr) > 2 else (tl[-2] if len(tl) > 2 else -1)
    bl = _sklansky(P, 0, miny, 1, -1)
    br = _sklansky(P, total - 1, miny, 1, 1)
    blc, brc = len(bl), len(br)
    if stop_idx >= 0:
        check_idx = bl[1] if blc > 2 else (br[2 - blc] if blc + brc > 2 else -1)
        if check_idx == stop_idx or (check_idx >= 0 and P[check_idx] == P[stop_idx]):
            # all points on one line: the bottom part mirrors the top part
            blc, brc = min(blc, 2), min(brc, 2)
    hull += [order[bl[i]] for i in range(blc - 1)]
    hull += [order[br[i]] for i in range(brc - 1, 0, -1)]
    nout = len(hull)
    if nout >= 3:
        min_i = max_i = lt = 0
        for i in range(1, nout):
            lt += hull[i - 1] < hull[i]
            if 1 < lt <= i - 2:
                break
            if hull[i] < hull[min_i]:
                min_i = i
            if hull[i] > hull[max_i]:
                max_i = i
        mmdist = abs(max_i - min_i)
        if (mmdist == 1 or mmdist == nout - 1) and (lt <= 1 or lt >= nout - 2):
            ascending = (max_i + 1) % nout == min_i
            i0 = min_i if ascending else max_i
            if i0 > 0:
                rot = hull[i0:] + hull[:i0]
                if all(ascending == (rot[i] < rot[i + 1]) for i in range(nout - 1)):
                    hull = rot
    return np.array([pts[k] for k in hull], dtype=np.int64)


def _cv_shift_start(pos: List[int]) -> int:
    """OpenCV's cyclic-shift test on the contour positions of the hull vertices (natural order):
    index the output starts at (0: no shift)."""
    nout = len(pos)
    if nout < 3:
        return 0
    min_i = max_i = lt = 0
    for i in range(1, nout):
        lt += pos[i - 1] < pos[i]
        if 1 < lt <= i - 2:
            break
        if pos[i] < pos[min_i]:
            min_i = i
        if pos[i] > pos[max_i]:
            max_i = i
    mmdist = abs(max_i - min_i)
    if (mmdist == 1 or mmdist == nout - 1) and (lt <= 1 or lt >= nout - 2):
        ascending = (max_i + 1) % nout == min_i
        i0 = min_i if ascending else max_i
        if i0 > 0:
            rot = pos[i0:] + pos[:i0]
            if all(ascending == (rot[i] < rot[i + 1]) for i in range(nout - 1)):
                return i0
    return 0


def hull_order_from_visits(contour_pts: np.ndarray) -> np.ndarray:
    """The same order as ``convex_hull_cv`` from what a border trace can carry along: the strict
    hull and, per hull vertex, the position of its FIRST and of its LAST visit in the
    CHAIN_APPROX_SIMPLE point list (the closed form of what OpenCV's sort + Sklansky walks do to a
    vertex the border passes more than once; the specification of the open device fix, DESIGN.md
    section 4).  Natural order: clockwise on screen from the (x, y)-largest vertex.  Position of a
    vertex: last visit for the (x, y)-largest vertex and for the inner vertices of the two LEFT
    chains (top -> (x, y)-smallest -> bottom); first visit for everything else (the (x, y)-smallest
    vertex, top = (max y, min x), bottom = (min y, min x), the right chains).  Then OpenCV's shift
    test.  Hulls of fewer than three vertices: as ``hull_like_cv``."""
    pts = [(int(x), int(y)) for x, y in np.asarray(contour_pts).reshape(-1, 2)]
    cyc = [tuple(int(v) for v in q) for q in hull_like_cv(np.asarray(pts))]
    n = len(cyc)
    if n < 3:
        return np.array(cyc, dtype=np.int64).reshape(-1, 2)
    first, last = {}, {}
    for i, q in enumerate(pts):
        first.setdefault(q, i)
        last[q] = i
    k = cyc.index(max(cyc))
    nat = cyc[k:] + cyc[:k]
    ymax, ymin = max(q[1] for q in nat), min(q[1] for q in nat)
    top = min(q for q in nat if q[1] == ymax)
    bot = min(q for q in nat if q[1] == ymin)
    it, im, ib = nat.index(top), nat.index(min(nat)), nat.index(bot)
    if im < it:
        im += n
    if ib < im:
        ib += n
    pos = []
    for j, q in enumerate(nat):
        left_inner = any(it < jj < im or im < jj < ib for jj in (j, j + n))
        use_last = j == 0 or (left_inner and q not in (top, bot, min(nat)))
        pos.append(last[q] if use_last else first[q])
    s0 = _cv_shift_start(pos)
    return np.array(nat[s0:] + nat[:s0], dtype=np.int64)


def hull_like_cv(contour_pts: np.ndarray) -> np.ndarray:
    """THE DEVICE'S RULE (csrc/contour_common.cuh::Hull): hull of a traced external contour
    clockwise on screen, cyclically shifted so that the contour's start pixel (raster-first
    pixel of the component, always a hull vertex) comes LAST.  This is the order
    cv2.convexHull(c, clockwise=False) returns for every SIMPLE contour (its hull indices then
    descend after OpenCV's cyclic shift); for a contour that visits a hull vertex twice OpenCV's
    shift test can fail -- ``convex_hull_cv`` is the exact restatement, and the two differ in
    the START of the cycle only, which matters to minAreaRect when two rectangles tie in float32
    area ("Known deviation")."""
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
