"""CPU tests: pin the oracle (oracle/) to the libraries the reference calls and to the
golden fixtures / known-answer vectors.  No GPU needed."""
import json
import math
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import contours_py as cp
from oracle import d2
from oracle import imutils_port
from oracle import measure as M
from oracle import pipeline as P
from uwcv import synth


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        return json.load(f)


# ---------------------------------------------------------------- paste ------------
def test_cpu_capability_is_vectorised():
    assert d2.assert_cpu_capability() in ("AVX2", "AVX512")


def test_scalar_recipe_equals_grid_sample():
    """The scalar float32 recipe the CUDA kernel implements == torch CPU grid_sample,
    value for value; detects a torch upgrade that changes the kernel."""
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    H, W = 96, 112
    total = 0
    for trial in range(40):
        m = torch.rand(28, 28)
        if trial % 4 == 1:
            m = (m > 0.5).float()
        if trial % 4 == 2:
            m = torch.full((28, 28), 0.5)
        if trial % 7 == 3:
            m = torch.ones(28, 28)
        x0, y0 = rng.random(2) * 100 - 10
        w, h = np.exp(rng.random(2) * 6 - 2)
        b = d2.Boxes(torch.tensor([[x0, y0, x0 + w, y0 + h]], dtype=torch.float32))
        b.clip((H, W))
        if not bool(b.nonempty()[0]):
            continue
        ref_val, _ = d2._do_paste_mask(m[None, None], b.tensor, H, W, skip_empty=False)
        ref_mask = d2.paste_masks_in_image(m[None], b.tensor, (H, W))[0].numpy()
        val, msk = d2.paste_scalar_recipe(m.numpy(), b.tensor[0].numpy(), np.arange(H), np.arange(W))
        rv = ref_val[0].numpy()
        assert np.array_equal(msk, ref_mask)
        assert np.array_equal(val, rv, equal_nan=True)
        total += H * W
    assert total > 200000


def test_paste_known_answers(kat):
    assert kat["paste_all_ones_count"] == 1600
    assert kat["paste_all_half_count"] == 1509
    ones = d2.paste_masks_in_image(torch.ones(1, 28, 28), torch.tensor([[10., 20., 50., 60.]]), (100, 100))[0]
    assert int(ones.sum()) == 1600 and bool(ones[20:60, 10:50].all())
    half = d2.paste_masks_in_image(torch.full((1, 28, 28), 0.5),
                                   torch.tensor([[10.5, 20.5, 50.5, 60.5]]), (100, 100))[0]
    assert int(half.sum()) == 1509


def test_tiny_and_degenerate_boxes():
    m = torch.rand(3, 28, 28)
    boxes = torch.tensor([[0., 0., 1e-30, 5.], [3., 4., 3.0 + 1e-3, 9.], [10.2, 10.2, 10.7, 30.]])
    out = d2.paste_masks_in_image(m, boxes, (40, 40))
    assert out.shape == (3, 40, 40) and out.dtype == torch.bool
    assert not torch.isnan(out.float()).any()


# ---------------------------------------------------------------- contours ---------
def _random_masks(rng, n):
    for t in range(n):
        h, w = rng.integers(3, 22, 2)
        p = rng.choice([0.2, 0.4, 0.5, 0.6, 0.8])
        yield (rng.random((h, w)) < p).astype(np.uint8)


def test_border_following_equals_cv2():
    rng = np.random.default_rng(0)
    masks = list(_random_masks(rng, 300))
    m = np.zeros((9, 9), np.uint8); m[1:8, 1:8] = 1; m[2:7, 2:7] = 0; m[4, 4] = 1
    masks.append(m)                                       # thin ring + island
    m = np.zeros((12, 12), np.uint8); m[1:11, 1:11] = 1; m[3:9, 3:9] = 0; m[5:7, 5:7] = 1
    masks.append(m)                                       # thick ring + island (nested: dropped)
    for mask in masks:
        cs, _ = cv2.findContours(mask.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        mine = cp.find_external_contours(mask)
        cs = list(cs)[::-1]                               # cv2 returns reverse raster order
        assert len(cs) == len(mine)
        for c, (pts, _start) in zip(cs, mine):
            assert np.array_equal(c.reshape(-1, 2), pts)
            assert cp.contour_area(pts) == cv2.contourArea(c)
            assert cp.arc_length(pts) == cv2.arcLength(c, True)


def _blob(rng, h, w, speckle):
    yy, xx = np.mgrid[:h, :w]
    m = np.zeros((h, w), bool)
    for _ in range(rng.integers(1, 3)):
        cx, cy = rng.random(2) * (w - 20) + 10
        a, b = rng.random(2) * 14 + 1
        th = rng.random() * np.pi
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        m |= (u / a) ** 2 + (v / b) ** 2 <= 1
    if speckle:
        m &= rng.random((h, w)) < 0.85
    return m.astype(np.uint8)


def test_min_area_rect_restatement_is_bit_exact():
    """hull (cv order) -> float32 rotating calipers -> rect -> boxPoints == cv2, bit for bit."""
    rng = np.random.default_rng(5)
    n = 0
    for t in range(250):
        m = _blob(rng, 64, 64, t % 3 == 0)
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for c in cs:
            r = cv2.minAreaRect(c)
            rr = cp.min_area_rect_cv(cp.hull_like_cv(c.reshape(-1, 2)))
            assert rr == ((r[0][0], r[0][1]), (r[1][0], r[1][1]), r[2])
            assert np.array_equal(cp.box_points_cv(rr), cv2.boxPoints(r))
            n += 1
    assert n > 250
    # speckle: contours that are NOT simple (one-pixel spurs, pixels visited twice) -- the exact
    # restatement of OpenCV's convexHull (order included) feeds the same calipers
    differ = 0
    for t in range(1200):
        h, w = rng.integers(6, 30), rng.integers(6, 30)
        m = (rng.random((h, w)) < rng.uniform(0.3, 0.8)).astype(np.uint8)
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        for c in cs:
            hull = cp.convex_hull_cv(c.reshape(-1, 2))
            assert np.array_equal(hull, cv2.convexHull(c).reshape(-1, 2))
            r = cv2.minAreaRect(c)
            rr = cp.min_area_rect_cv(hull)
            assert rr == ((r[0][0], r[0][1]), (r[1][0], r[1][1]), r[2])
            assert np.array_equal(cp.box_points_cv(rr), cv2.boxPoints(r))
            if len(hull) >= 3:        # the closed form a border trace can carry (first / last visits)
                assert np.array_equal(cp.hull_order_from_visits(c.reshape(-1, 2)), hull)
            simple = len({tuple(q) for q in c.reshape(-1, 2).tolist()}) == len(c)
            same = np.array_equal(cp.hull_like_cv(c.reshape(-1, 2)), hull)
            assert same or not simple or len(hull) <= 2      # the device's rule: every simple contour
            differ += not same
            n += 1
    assert n > 5000 and differ > 50          # (the sample does contain what the rule misses)
    for pts in ([[2, 3], [7, 3]], [[2, 3], [2, 9]], [[1, 1], [5, 5]], [[5, 1], [1, 5]], [[3, 3]],
                [[0, 0], [9, 2]], [[4, 1], [5, 9]]):
        c = np.array(pts, np.int32).reshape(-1, 1, 2)
        r = cv2.minAreaRect(c)
        rr = cp.min_area_rect_cv(cp.hull_like_cv(c.reshape(-1, 2)))
        assert rr == ((r[0][0], r[0][1]), (r[1][0], r[1][1]), r[2])
        assert np.array_equal(cp.box_points_cv(rr), cv2.boxPoints(r))


def test_known_deviation_hull_start_of_a_non_simple_contour():
    """DESIGN.md section 4, "Known deviation": the one instance of the 77 000-seed sweep whose
    min-area rectangle differs on the device.  The contour has a one-pixel diagonal spur (hull
    vertex (495, 407) is visited twice), cv2.convexHull therefore does NOT rotate its output
    behind the contour's start pixel, and two rectangles of equal float32 area (35.0) swap places
    in the calipers' last-minimum rule.  The restatement below follows the device code ("start
    pixel last"): this test pins what differs (the hull's START only) and what does not (the hull
    itself, the tie), so a fix has to change exactly this."""
    pts = [[496, 402], [497, 403], [497, 404], [498, 405], [497, 406], [496, 406], [495, 407], [494, 406],
           [495, 407], [496, 406], [497, 407], [500, 407], [501, 406], [500, 405], [500, 404], [501, 403],
           [500, 404], [500, 405], [499, 406], [497, 404], [497, 403]]
    c = np.array(pts, np.int32).reshape(-1, 1, 2)
    hull_cv = cv2.convexHull(c).reshape(-1, 2)
    hull_dev = cp.hull_like_cv(c.reshape(-1, 2))
    assert hull_cv.tolist() == [[501, 406], [500, 407], [495, 407], [494, 406], [496, 402], [501, 403]]
    assert np.array_equal(np.roll(hull_cv, 1, axis=0), hull_dev)          # same polygon, other start
    # the exact restatement of OpenCV's convexHull (contour positions of the hull vertices
    # 12, 11, 6, 7, 0, 15: 6 < 7, not a rotation of a descending sequence -> no cyclic shift)
    assert np.array_equal(cp.convex_hull_cv(c.reshape(-1, 2)), hull_cv)
    r_cv = cv2.minAreaRect(c)
    r_dev = cp.min_area_rect_cv(hull_dev)
    assert r_cv == ((497.5, 404.5), (5.0, 7.0), -90.0)
    assert r_dev != r_cv and abs(r_dev[2] + 78.69007) < 1e-4
    assert abs(r_dev[1][0] * r_dev[1][1] - 35.0) < 1e-4                   # the tie
    # started where OpenCV starts, the same calipers give OpenCV's rectangle
    assert cp.min_area_rect_cv(hull_cv) == r_cv
    assert np.array_equal(cp.hull_order_from_visits(c.reshape(-1, 2)), hull_cv)


def test_contour_known_answers(kat):
    r = kat["rect_10x25"][0]
    assert r["points"] == [[5, 10], [5, 19], [29, 19], [29, 10]]
    assert r["area"] == 216.0 and r["arclen"] == 66.0
    assert r["rect"][:4] == [17.0, 14.5, 9.0, 24.0]
    d = kat["disc_r40"]
    assert d["pixels"] == 5025 and d["contours"][0]["area"] == 4912.0
    assert abs(d["contours"][0]["arclen"] - 263.7645) < 1e-4
    assert abs(d["contours"][0]["rect"][2] - 79.19595) < 1e-4
    assert kat["single_pixel"][0]["points"] == [[3, 3]] and kat["single_pixel"][0]["arclen"] == 0.0
    assert kat["line_1x6"][0]["area"] == 0.0 and kat["line_1x6"][0]["arclen"] == 10.0
    assert kat["square_spur"][0]["area"] == 10.0
    assert abs(kat["square_spur"][0]["arclen"] - 16.828427) < 1e-5
    assert len(kat["ring"]) == 1 and kat["ring"][0]["area"] == 49.0 and kat["ring"][0]["arclen"] == 28.0
    assert len(kat["diag_squares"]) == 1 and kat["diag_squares"][0]["area"] == 2.0
    # the pure-Python restatement reproduces every stored contour
    cases = {
        "rect_10x25": ((30, 40), lambda m: m.__setitem__((slice(10, 20), slice(5, 30)), 1)),
    }
    m = np.zeros((30, 40), np.uint8); m[10:20, 5:30] = 1
    assert cp.find_external_contours(m)[0][0].tolist() == r["points"]


def test_union_known_answer(kat):
    """Reference-literal GetMask_Contours on three painted ellipses (SURVEY.md 8(c) iv)."""
    survey = [
        (59.116913, 2.077556, 0.481335, 0.602998, 0.776530, 28.455027, 59.116913, 33.396974, 124.568542),
        (92.756706, 1.0, 1.0, 0.754144, 0.868415, 92.756706, 92.756706, 79.083201, 263.764500),
        (70.832909, 3.006643, 0.332597, 0.483796, 0.695555, 23.558804, 70.832909, 33.358827, 138.911687),
    ]
    yy, xx = np.mgrid[:200, :200]
    masks = np.stack([((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1
                      for cx, cy, a, b in ((100, 90, 40, 40), (40, 40, 25, 12), (160, 160, 10, 30))])
    for literal in (False, True):
        rows = M.get_mask_contours((200, 200, 3), np.array([0, 0, 0]), masks, [0], literal_paint=literal)
        assert np.allclose(rows, np.array(survey), rtol=0, atol=2e-6)
        assert np.array_equal(rows, np.array(kat["union_three_ellipses"]))
    assert M.get_mask_contours((200, 200, 3), np.array([1, 1, 1]), masks, [0]) is None
    with pytest.raises(ValueError):                      # masks exist but are all-false
        M.get_mask_contours((200, 200, 3), np.array([0]), np.zeros((1, 200, 200), bool), [0])


def test_imutils_port():
    pts = np.array([[10, 0], [0, 5], [10, 5], [0, 0]])
    o = imutils_port.order_points(pts)
    assert o.dtype == np.float32
    assert o.tolist() == [[0, 0], [10, 0], [10, 5], [0, 5]]
    assert imutils_port.grab_contours((["a"], "h")) == ["a"]
    assert imutils_port.grab_contours(("img", ["a"], "h")) == ["a"]
    with pytest.raises(Exception):
        imutils_port.grab_contours((1,))


def test_frame_moments_equals_cv2_on_padded_frame():
    rng = np.random.default_rng(0)
    for t in range(60):
        H, W = 300, 400
        h, w = rng.integers(1, 60, 2)
        y0 = rng.integers(0, H - h); x0 = rng.integers(0, W - w)
        win = (rng.random((h, w)) < rng.choice([0.1, 0.5, 0.9])).astype(np.uint8)
        full = np.zeros((H, W), np.uint8); full[y0:y0 + h, x0:x0 + w] = win
        a = cv2.moments(full, binaryImage=True)
        b = M.frame_moments(win, x0, y0)
        for k, v in b.items():
            assert a[k] == v, k


def test_get_counts_quirks():
    c = M.get_counts(np.array([0, 1, 1, 2, 3, 3, 3]))
    assert c["SCount"] == 2 and c["WTCount"] == 1 and c["PTCount"] == 3 and c["PCount"] == 3
    assert c["TotalCount"] == 6 and c["intended"] == [1, 2, 1, 3]


def test_moving_average_and_report():
    assert M.moving_average([1, 2, 3, 4.005], 3) == [2.0, round((2 + 3 + 4.005) / 3, 2)]
    assert M.moving_average([1, 2], 3) == []
    rows = np.arange(45, dtype=np.float64).reshape(5, 9)
    sm, h = M.report_class(rows)
    assert sm.shape == (3, 9) and set(h) == set(M.CSV_COLUMNS)


# ---------------------------------------------------------------- NMS --------------
def _nms_numpy(boxes, scores, cls, thr):
    """float32 restatement the CUDA kernel implements."""
    f32 = np.float32
    b = boxes.astype(f32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    keep = []
    dead = np.zeros(len(b), bool)
    for i in order:
        if dead[i]:
            continue
        keep.append(i)
        w = np.maximum(f32(0), np.minimum(b[i, 2], b[:, 2]) - np.maximum(b[i, 0], b[:, 0]))
        h = np.maximum(f32(0), np.minimum(b[i, 3], b[:, 3]) - np.maximum(b[i, 1], b[:, 1]))
        inter = w * h
        with np.errstate(all="ignore"):
            iou = inter / ((area[i] + area) - inter)
        dead |= (iou.astype(np.float64) > thr) & (cls == cls[i])
    return np.array(keep)


def test_nms_oracle_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_small.npz"))
    b, s, c = torch.from_numpy(g["boxes"]), torch.from_numpy(g["scores"]), torch.from_numpy(g["classes"])
    keep = d2.batched_nms_vanilla(b, s, c, 0.5).numpy()
    assert np.array_equal(keep, g["keep_all"])
    assert np.array_equal(_nms_numpy(g["boxes"], g["scores"], g["classes"], 0.5), keep)
    assert len(np.unique(g["scores"])) == len(g["scores"])


def test_fast_rcnn_inference_oracle_shapes():
    torch.manual_seed(0)
    R, K = 50, 4
    boxes = torch.rand(R, K, 2) * 80
    boxes = torch.cat([boxes, boxes + torch.rand(R, K, 2) * 40 + 1], dim=2).reshape(R, K * 4)
    scores = torch.rand(R, K + 1)
    res, rows = d2.fast_rcnn_inference_single_image(boxes, scores, (100, 100), 0.5, 0.5, 20)
    assert len(res) <= 20 and len(rows) == len(res)
    assert bool((res.scores[:-1] >= res.scores[1:]).all())


# ---------------------------------------------------------------- golden rows ------
def test_golden_blobs_reproduce(golden_dir):
    g = np.load(os.path.join(golden_dir, "blobs_small.npz"))
    H, W = int(g["H"]), int(g["W"])
    batch = [synth.blob_instances(k, 40, 256, 333, seed=77, size_range=(4.0, 90.0)) for k in range(3)]
    ri, rf = P.oracle_table(batch, (H, W))
    assert np.array_equal(ri, g["rows_i"])
    assert np.array_equal(rf, g["rows_f"], equal_nan=True)
    assert ri.shape[1] == len(M.INT_COLUMNS) and rf.shape[1] == len(M.FLOAT_COLUMNS)


def test_golden_c1_consistency(golden_dir):
    g = np.load(os.path.join(golden_dir, "c1_maskrcnn.npz"))
    assert g["masks"].shape == (200, 1, 28, 28)
    ri = g["rows_i"]
    area = ri[:, M.INT_COLUMNS.index("area_px")]
    valid = ri[:, M.INT_COLUMNS.index("valid")]
    assert np.array_equal(valid == 1, area > 0)
    # spot-check three instances end to end against the stored rows
    from uwcv.structures import Boxes, Instances
    inst = Instances((1024, 1024))
    sel = [int(np.argmax(area)), int(np.flatnonzero(valid)[0]), int(np.flatnonzero(valid == 0)[0])]
    inst.pred_boxes = Boxes(torch.from_numpy(g["boxes"][sel]))
    inst.scores = torch.from_numpy(g["scores"][sel])
    inst.pred_classes = torch.from_numpy(g["classes"][sel])
    inst.pred_masks = torch.from_numpy(g["masks"][sel])
    ri2, rf2 = P.oracle_table([inst], (1024, 1024))
    skip = M.INT_COLUMNS.index("inst_idx")
    cols = [j for j in range(ri.shape[1]) if j != skip]
    assert np.array_equal(ri2[:, cols], ri[sel][:, cols])
    assert np.array_equal(rf2, g["rows_f"][sel], equal_nan=True)


def test_mask_rcnn_inference_restatement():
    """detectron2 mask_head.py::mask_rcnn_inference: predicted-class channel, sigmoid, split per
    image (class-agnostic head: channel 0)."""
    from oracle import d2
    g = torch.Generator().manual_seed(3)
    logits = torch.randn((7, 4, 28, 28), generator=g) * 4
    cls = torch.tensor([0, 3, 1, 2, 2, 0, 1])
    insts = [d2.Instances((50, 60), pred_classes=cls[:3]), d2.Instances((50, 60), pred_classes=cls[3:])]
    d2.mask_rcnn_inference(logits, insts)
    assert insts[0].pred_masks.shape == (3, 1, 28, 28) and insts[1].pred_masks.shape == (4, 1, 28, 28)
    for k in range(7):
        got = (insts[0] if k < 3 else insts[1]).pred_masks[k if k < 3 else k - 3, 0]
        # torch's CPU sigmoid is not bit-stable across tensor shapes (vectorised body vs scalar
        # tail differ by an ulp), which is why the f4 parity bar is torch's CUDA sigmoid
        assert torch.allclose(got, torch.sigmoid(logits[k, cls[k]]), rtol=3e-7, atol=0)
    ag = [d2.Instances((50, 60), pred_classes=cls)]
    d2.mask_rcnn_inference(logits[:, :1], ag)
    assert torch.equal(ag[0].pred_masks, torch.sigmoid(logits[:, :1]))
