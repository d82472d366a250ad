"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bar: bit-exact for masks, pixel areas, bboxes, raw moments, contour counts / points and
instance counts; float columns within 1e-5 relative (most are expected bit-exact, the
test reports how many are)."""
import os
import zlib

import numpy as np
import pytest
import torch

import uwcv
from uwcv import api, schema, synth
from oracle import d2, measure as M, pipeline as P

pytestmark = pytest.mark.gpu

RTOL = 1e-5      # north_star tolerance for float measurements
IC, FC = schema.ICOL, schema.FCOL


from oracle.compare import compare_tables  # noqa: E402,F401  (the per-column parity rule)


def plane_crcs(planes):
    p = planes.cpu().numpy().view(np.uint32)
    return np.array([zlib.crc32(p[i].tobytes()) for i in range(p.shape[0])], dtype=np.int64)


# ---------------------------------------------------------------------------------------
def test_golden_blobs_rows_and_planes(golden_dir):
    g = np.load(os.path.join(golden_dir, "blobs_small.npz"))
    H, W = int(g["H"]), int(g["W"])
    batch = [synth.blob_instances(k, 40, 256, 333, seed=77, size_range=(4.0, 90.0)) for k in range(3)]
    table, planes = uwcv.measure_instances(batch, (H, W), return_planes=True)
    exact = compare_tables(table, g["rows_i"], g["rows_f"])
    assert np.array_equal(plane_crcs(planes), g["plane_crc"])
    print(f"blobs_small: {exact}/{schema.NUM_FLOAT} float columns bit-exact")
    assert exact >= 24


def test_golden_c1_maskrcnn(golden_dir):
    """Config 1: raw random-init Mask R-CNN head output (sub-pixel boxes, saturated
    probabilities, empty and multi-component masks)."""
    g = np.load(os.path.join(golden_dir, "c1_maskrcnn.npz"))
    inst = uwcv.Instances((1024, 1024))
    inst.pred_boxes = uwcv.Boxes(torch.from_numpy(g["boxes"]))
    inst.scores = torch.from_numpy(g["scores"])
    inst.pred_classes = torch.from_numpy(g["classes"])
    inst.pred_masks = torch.from_numpy(g["masks"])
    table, planes = uwcv.measure_instances(inst, (1024, 1024), return_planes=True)
    assert len(table) == 200
    exact = compare_tables(table, g["rows_i"], g["rows_f"])
    assert np.array_equal(plane_crcs(planes), g["plane_crc"])
    print(f"c1: {exact}/{schema.NUM_FLOAT} float columns bit-exact; "
          f"{int((table['valid'] == 0).sum())} empty masks, max contours {table['n_contours'].max()}")
    # cropped mode (no planes) gives the same rows
    t2 = uwcv.measure_instances(inst, (1024, 1024))
    assert np.array_equal(t2.ints, table.ints) and np.array_equal(t2.floats, table.floats, equal_nan=True)


def test_paste_drop_in_equals_detectron2_restatement():
    torch.manual_seed(3)
    H, W = 200, 333                                       # W not a multiple of 32
    n = 60
    inst = synth.blob_instances(5, n, H, W, seed=11, size_range=(2.0, 150.0))
    masks = inst.pred_masks[:, 0]
    boxes = inst.pred_boxes.tensor
    ref = d2.paste_masks_in_image(masks, boxes, (H, W))
    out = uwcv.paste_masks_in_image(masks, boxes, (H, W))
    assert out.dtype == torch.bool and out.is_cuda and tuple(out.shape) == (len(masks), H, W)
    assert torch.equal(out.cpu(), ref)
    # other thresholds
    for thr in (0.3, 0.7):
        assert torch.equal(uwcv.paste_masks_in_image(masks, boxes, (H, W), thr).cpu(),
                           d2.paste_masks_in_image(masks, boxes, (H, W), thr))
    # known answers
    ones = uwcv.paste_masks_in_image(torch.ones(1, 28, 28), torch.tensor([[10., 20., 50., 60.]]), (100, 100))
    assert int(ones.sum()) == 1600 and bool(ones[0, 20:60, 10:50].all())
    half = uwcv.paste_masks_in_image(torch.full((1, 28, 28), 0.5),
                                     torch.tensor([[10.5, 20.5, 50.5, 60.5]]), (100, 100))
    assert int(half.sum()) == 1509
    assert uwcv.paste_masks_in_image(torch.zeros(0, 28, 28), torch.zeros(0, 4), (8, 8)).shape == (0, 8, 8)
    with pytest.raises(ValueError):
        uwcv.paste_masks_in_image(masks, boxes, (H, W), threshold=0.0)


def test_paste_edge_boxes():
    """Sub-pixel, border-clipped, frame-filling and huge-coordinate boxes; saturated maps."""
    H, W = 96, 160
    boxes = torch.tensor([
        [0., 0., 1e-30, 5.],            # positive-but-tiny width
        [3., 4., 3.001, 9.],
        [10.2, 10.2, 10.7, 30.],        # 0.5 px wide
        [0., 0., 160., 96.],            # whole frame
        [150.5, 80.5, 160., 96.],       # touches the bottom-right corner
        [0., 0., 7.3, 2.2],
        [31.5, 0., 32.5, 96.],          # straddles a word boundary
        [63.9, 10., 96.1, 50.],
        [20., 95.4, 100., 96.],
    ])
    g = torch.Generator().manual_seed(0)
    masks = torch.rand(len(boxes), 28, 28, generator=g)
    masks[3] = (masks[3] > 0.5).float()
    masks[4] = 1.0
    masks[7] = 0.5
    ref = d2.paste_masks_in_image(masks, boxes, (H, W))
    out = uwcv.paste_masks_in_image(masks, boxes, (H, W)).cpu()
    assert torch.equal(out, ref)


def test_detector_postprocess_drop_in():
    inst = synth.blob_instances(2, 50, 256, 333, seed=5, size_range=(3.0, 120.0))
    inst.pred_boxes.tensor[3] = torch.tensor([10., 10., 10., 40.])      # empty -> dropped
    ref = d2.detector_postprocess(P.to_oracle_instances(inst), 320, 416)
    out = uwcv.detector_postprocess(inst, 320, 416)
    assert len(out) == len(ref) == 49
    assert out.image_size == (320, 416)
    assert torch.equal(out.pred_boxes.tensor.cpu(), ref.pred_boxes.tensor)
    assert torch.equal(out.pred_masks.cpu(), ref.pred_masks)
    assert torch.equal(out.pred_classes.cpu(), ref.pred_classes)
    assert torch.equal(out._fields["scores"].cpu(), ref.scores)


def test_measure_classes_of_interest_and_empty():
    batch = [synth.blob_instances(k, 30, 200, 200, seed=21) for k in range(2)]
    full = uwcv.measure_instances(batch, (200, 200))
    for cls in range(4):
        t = uwcv.measure_instances(batch, (200, 200), classes_of_interest=[cls])
        sel = full.for_class(cls)
        skip = IC["inst_idx"]
        cols = [j for j in range(schema.NUM_INT) if j != skip]
        assert np.array_equal(t.ints[:, cols], sel.ints[:, cols])
        assert np.array_equal(t.floats, sel.floats, equal_nan=True)
        ri, rf = P.oracle_table(batch, (200, 200), classes_of_interest=[cls])
        compare_tables(t, ri, rf)
    assert len(uwcv.measure_instances(batch, (200, 200), classes_of_interest=[9])) == 0
    assert len(uwcv.measure_instances([], (200, 200))) == 0
    # instance counts per class == what GetCounts intends
    counts = [sum(uwcv.get_counts(b)[k] for b in batch) for k in range(4)]
    assert counts == [int((full["class_id"] == k).sum()) for k in range(4)]


def test_reference_literal_union_rows_for_disjoint_instances():
    """When instances of a class do not touch, the reference's union-contour rows
    (GetMask_Contours) are exactly the per-instance rows sorted left to right."""
    H = W = 400
    n = 9
    inst = uwcv.Instances((H, W))
    cx = torch.tensor([50., 150., 250.]).repeat(3) + torch.tensor([0., 7., 13.]).repeat_interleave(3)
    cy = torch.tensor([60., 180., 300.]).repeat_interleave(3)
    half = torch.linspace(22, 40, n)
    inst.pred_boxes = uwcv.Boxes(torch.stack([cx - half, cy - half * 0.8, cx + half, cy + half * 0.8], 1))
    inst.scores = torch.linspace(0.9, 0.5, n)
    inst.pred_classes = torch.zeros(n, dtype=torch.int64)
    lin = (torch.arange(28, dtype=torch.float32) + 0.5) / 28 * 2 - 1
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    ang = torch.linspace(0, 2.5, n)[:, None, None]
    u = xx * torch.cos(ang) + yy * torch.sin(ang)
    v = -xx * torch.sin(ang) + yy * torch.cos(ang)
    inst.pred_masks = torch.sigmoid(9 * (0.8 - torch.sqrt((u / 0.9) ** 2 + (v / 0.55) ** 2)))[:, None]
    rows_ref = P.reference_literal_rows(inst, (H, W), [0])
    t = uwcv.measure_instances(inst, (H, W), classes_of_interest=[0])
    assert (t["n_contours"] == 1).all()
    mine = t.reference_rows()
    assert mine.shape == rows_ref.shape
    # the reference orders rows by boundingRect x (stable over cv2's reverse-raster order);
    # compare as multisets of rows so that ties in x cannot matter
    mine = mine[np.lexsort(mine.T[::-1])]
    rows_ref = rows_ref[np.lexsort(rows_ref.T[::-1])]
    rel = np.abs(mine - rows_ref) / np.maximum(np.abs(rows_ref), 1e-30)
    assert rel.max() <= 1e-6, f"max rel err per column {rel.max(axis=0)}"
    print("union == per-instance rows; exact columns:", int((rel.max(axis=0) == 0).sum()), "/ 9")


def test_workspace_overflow_is_reported():
    eng = api.Engine.get()
    inst = synth.blob_instances(0, 20, 128, 128, seed=3)
    b, keep = api.scale_clip_boxes(inst.pred_boxes.tensor, (128, 128), (128, 128))
    dev = eng.device
    with pytest.raises(RuntimeError, match="tile words"):
        old = (eng._ws, eng._cap_n, eng._cap_words)
        try:
            eng._ws = torch.empty(eng.L.uwcv_workspace_bytes(20, 8), dtype=torch.uint8, device=dev)
            eng._cap_n, eng._cap_words = 20, 10 ** 9          # pretend it is big enough
            eng.run(inst.pred_masks[:, 0].contiguous().to(dev), b.contiguous().to(dev), 128, 128,
                    n_tile_words=8)
            eng.check_status()
        finally:
            eng._ws, eng._cap_n, eng._cap_words = old


# ---------------------------------------------------------------- NMS --------------
def test_nms_golden_and_batched(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_small.npz"))
    eng = api.Engine.get()
    dev = eng.device
    b = torch.from_numpy(g["boxes"]).to(dev)
    s = torch.from_numpy(g["scores"]).to(dev)
    c = torch.from_numpy(g["classes"]).to(dev)
    R = len(b)
    keep, cnt = eng.nms(b, s, c, [0, R], 0.05, 0.5, -1)
    k = int(cnt[0])
    assert np.array_equal(keep[:k].cpu().numpy(), g["keep_thr005"])
    keep, cnt = eng.nms(b, s, c, [0, R], -1.0, 0.5, 100)
    assert int(cnt[0]) == 100 and np.array_equal(keep[:100].cpu().numpy(), g["keep_all"][:100])
    # three images in one call (one of them empty), each equals its own single call
    off = [0, 700, 700, R]
    keep, cnt = eng.nms(b, s, c, off, 0.05, 0.5, -1)
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    assert cnt[1] == 0
    for lo, hi, n in ((0, 700, cnt[0]), (700, R, cnt[2])):
        bb, ss, cc = b[lo:hi].cpu(), s[lo:hi].cpu(), c[lo:hi].cpu()
        ref = d2.batched_nms_vanilla(bb, ss, cc, 0.5)
        ref = ref[ss[ref] > 0.05].numpy() + lo
        assert np.array_equal(keep[lo:lo + n], ref)


def test_nms_dense_config4_style():
    """~20 k clustered candidates, 4 classes: keep list == vanilla."""
    b, s, c = synth.clustered_candidates(5000, 4096, 4096, seed=99)
    ref = d2.batched_nms_vanilla(b, s, c, 0.5)
    ref = ref[s[ref] > 0.05].numpy()
    eng = api.Engine.get()
    dev = eng.device
    keep, cnt = eng.nms(b.to(dev), s.to(dev), c.to(dev), [0, len(b)], 0.05, 0.5, 6000)
    k = int(cnt[0])
    assert np.array_equal(keep[:k].cpu().numpy(), ref[:6000])
    print(f"dense NMS: {len(b)} candidates -> {k} kept")


def test_fast_rcnn_inference_drop_in():
    torch.manual_seed(0)
    R, K = 300, 4
    xy = torch.rand(R, K, 2) * 300
    boxes = torch.cat([xy, xy + torch.rand(R, K, 2) * 80 + 1], dim=2).reshape(R, K * 4)
    scores = torch.rand(R, K + 1)
    boxes[7, 3] = float("inf")                            # non-finite proposal is dropped
    ref, ref_rows = d2.fast_rcnn_inference_single_image(boxes, scores, (320, 320), 0.8, 0.5, 100)
    out, rows = uwcv.fast_rcnn_inference_single_image(boxes, scores, (320, 320), 0.8, 0.5, 100)
    assert len(out) == len(ref)
    assert torch.equal(out.pred_boxes.tensor.cpu(), ref.pred_boxes.tensor)
    assert torch.equal(out.scores.cpu(), ref.scores)
    assert torch.equal(out.pred_classes.cpu(), ref.pred_classes)
    # kept proposal rows index the finite-filtered list, as Detectron2's filter_inds[:, 0] do
    assert torch.equal(rows.cpu(), ref_rows)


# ---------------------------------------------------------------- full size --------
def test_full_size_properties_and_sampled_parity():
    """BASELINE config-2 shape (2048 x 2048, 1000 instances / image; 4 images here):
    size-independent properties on every instance + oracle parity on a sample."""
    H = W = 2048
    batch = synth.blob_batch(4, 1000, H, W, seed=1234)
    table, planes = uwcv.measure_instances(batch, (H, W), return_planes=True)
    n = len(table)
    assert n == sum(len(b) for b in batch)
    p = planes.view(torch.int32)
    # (1) popcount of every plane == area_px (bits written nowhere else, nothing lost)
    pc = torch.zeros(n, dtype=torch.int64, device=p.device)
    for lo in range(0, n, 250):
        x = p[lo:lo + 250].to(torch.int64) & 0xFFFFFFFF
        cnt = torch.zeros_like(x)
        for sh in range(32):
            cnt += (x >> sh) & 1
        pc[lo:lo + 250] = cnt.sum(dim=(1, 2))
    assert np.array_equal(pc.cpu().numpy(), table["area_px"])
    # (2) bbox encloses the centroid, lies inside the clipped box (+1 px), m00 consistency
    v = table["valid"] == 1
    assert (table["bbox_x0"][v] <= table["cx"][v]).all() and (table["cx"][v] <= table["bbox_x1"][v]).all()
    assert (table["bbox_y0"][v] <= table["cy"][v]).all() and (table["cy"][v] <= table["bbox_y1"][v]).all()
    assert (table["m10"][v] >= table["bbox_x0"][v] * table["area_px"][v]).all()
    assert (table["contour_area"][v] <= table["area_px"][v]).all()
    assert (table["n_contours"][v] >= 1).all() and (table["n_contours"][~v] == 0).all()
    # (3) linearity: planes of image 0 unpacked and OR-ed == union popcount bound
    # (4) sampled oracle parity (every 23rd instance of each image)
    for k, inst in enumerate(batch):
        idx = torch.arange(0, len(inst), 23)
        sub = inst[idx]
        ri, rf = P.oracle_table([sub], (H, W), image_idx_offset=k)
        rows = np.flatnonzero(table["image_idx"] == k)[idx.numpy()]
        t = table.select(rows)
        compare_tables(t, ri, rf, skip_int=("inst_idx",))


def test_random_box_fuzz_against_grid_sample():
    """400 random boxes (extreme aspect ratios, sub-pixel, off-frame, frame-sized) with
    random / binary / constant probability maps: bit-exact against the oracle paste."""
    g = torch.Generator().manual_seed(2024)
    H, W = 144, 208
    n = 400
    cx = torch.rand(n, generator=g) * (W + 40) - 20
    cy = torch.rand(n, generator=g) * (H + 40) - 20
    w = torch.exp(torch.rand(n, generator=g) * 9 - 3)          # 0.05 .. 400 px
    h = torch.exp(torch.rand(n, generator=g) * 9 - 3)
    boxes = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    boxes[:, 0::2] = boxes[:, 0::2].clamp(0, W)
    boxes[:, 1::2] = boxes[:, 1::2].clamp(0, H)
    keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
    boxes = boxes[keep].contiguous()
    n = len(boxes)
    masks = torch.rand(n, 28, 28, generator=g)
    masks[::5] = (masks[::5] > 0.5).float()
    masks[1::7] = 0.5
    masks[2::11] = 1.0
    masks[3::13] = 0.0
    ref = d2.paste_masks_in_image(masks, boxes, (H, W))
    out = uwcv.paste_masks_in_image(masks, boxes, (H, W)).cpu()
    bad = (out != ref).flatten(1).any(1).nonzero().flatten()
    assert bad.numel() == 0, f"instances with differing masks: {bad[:10].tolist()}, boxes {boxes[bad[:3]]}"
    assert n > 300


def test_config4_dense_pipeline():
    """BASELINE config-4 shape: 4096 x 4096 micrograph, ~20 k clustered candidates ->
    score filter + per-class NMS on the GPU -> ~5-6 k small overlapping survivors ->
    paste + measure (full-frame planes would need 10.5 GB: cropped contract here) ->
    properties on all rows + oracle parity on a sample."""
    H = W = 4096
    b, s, c = synth.clustered_candidates(5000, H, W, seed=99)
    eng = api.Engine.get()
    dev = eng.device
    keep, cnt = eng.nms(b.to(dev), s.to(dev), c.to(dev), [0, len(b)], 0.05, 0.5, 6000)
    k = int(cnt[0])
    keep = keep[:k].cpu()
    ref = d2.batched_nms_vanilla(b, s, c, 0.5)
    ref = ref[s[ref] > 0.05][:6000]
    assert torch.equal(keep, ref)
    g = torch.Generator().manual_seed(8)
    inst = uwcv.Instances((H, W))
    inst.pred_boxes = uwcv.Boxes(b[keep])
    inst.scores = s[keep]
    inst.pred_classes = c[keep]
    inst.pred_masks = synth.blob_probs(k, g)[:, None]
    table = uwcv.measure_instances(inst, (H, W))
    assert len(table) == k and k > 4000
    v = table["valid"] == 1
    assert v.mean() > 0.95
    assert (table["bbox_x1"][v] < W).all() and (table["bbox_y1"][v] < H).all()
    assert (table["area_px"][v] >= table["contour_area"][v]).all()
    counts = uwcv.get_counts(inst)
    recs = uwcv.group_by_class(table)
    assert [r["count"] for r in recs] == counts
    idx = torch.arange(0, k, 41)
    ri, rf = P.oracle_table([inst[idx]], (H, W))
    compare_tables(table.select(idx.numpy()), ri, rf, skip_int=("inst_idx",))
    print(f"config 4: {len(b)} candidates -> {k} instances, {int(v.sum())} non-empty")


def test_union_mode_equals_reference_literal_rows():
    """f1: reference-literal GetMask_Contours (union of the class masks, every external
    contour, left-to-right) for overlapping / touching / multi-blob instances."""
    H, W = 384, 512
    batch = []
    for k in range(3):
        inst = synth.blob_instances(k, 120, H, W, seed=300, size_range=(10.0, 110.0))
        batch.append(inst)
    total = 0
    for cls in range(4):
        ut = uwcv.measure_union(batch, (H, W), classes_of_interest=[cls])
        for k, inst in enumerate(batch):
            try:
                ref = P.reference_literal_rows(inst, (H, W), [cls])
            except ValueError:
                ref = np.zeros((0, 9))
            ref = np.zeros((0, 9)) if ref is None else ref
            mine = ut.reference_rows(image_idx=k)
            assert mine.shape == ref.shape, (cls, k, mine.shape, ref.shape)
            if len(ref):
                rel = np.abs(mine - ref) / np.maximum(np.abs(ref), 1e-30)
                assert rel.max() <= 1e-6, (cls, k, rel.max(axis=0))
            total += len(ref)
    assert total > 60
    # union over all classes at once (classes_of_interest=None == every class)
    ut = uwcv.measure_union(batch[0], (H, W))
    ref = P.reference_literal_rows(batch[0], (H, W), [0, 1, 2, 3])
    assert np.allclose(ut.reference_rows(), ref, rtol=1e-6, atol=0)
    print(f"union mode: {total} reference rows matched")


# ---------------------------------------------------------------- stream (e2e form) ----
def test_measurement_stream_equals_synchronous_calls():
    """uwcv.MeasurementStream (two calls in flight, per-slot buffers) returns, in order, the
    tables of the synchronous call -- including when the cached workspace overflows and a
    call is transparently repeated, and for device-resident inputs."""
    H = W = 256
    batches = [[synth.blob_instances(4 * s + k, 20 + 7 * s + k, H, W, seed=50 + s) for k in range(3)]
               for s in range(5)]
    batches.append([synth.blob_instances(99, 400, H, W, seed=77) for _ in range(2)])  # overflow
    batches.append([])                                                                # empty
    for b in batches[:2]:                                                             # pinned inputs
        for inst in b:
            for k, v in list(inst.get_fields().items()):
                inst.set(k, v.pin_memory() if isinstance(v, torch.Tensor)
                         else uwcv.Boxes(v.tensor.pin_memory()))
    for inst in batches[2]:                                                           # device inputs
        for k, v in list(inst.get_fields().items()):
            inst.set(k, v.cuda() if isinstance(v, torch.Tensor) else uwcv.Boxes(v.tensor.cuda()))
    want = [uwcv.measure_instances(b, (H, W)) for b in batches]
    api.Engine._engines.clear()                      # fresh engine: small cached workspace
    for depth in (1, 2, 3):
        stream = uwcv.MeasurementStream(depth=depth)
        got = list(stream.map(batches, (H, W)))
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert np.array_equal(g.ints, w.ints)
            assert np.array_equal(g.floats, w.floats, equal_nan=True)
    # handles collected out of order
    stream = uwcv.MeasurementStream(depth=2)
    p0 = stream.submit(batches[0], (H, W))
    p1 = stream.submit(batches[1], (H, W))
    p2 = stream.submit(batches[3], (H, W))           # reuses slot 0: p0 is materialised first
    for p, w in ((p2, want[3]), (p1, want[1]), (p0, want[0])):
        assert np.array_equal(p.result().ints, w.ints)


# ---------------------------------------------------------------- single forward (f4) ---
def _logit_instances(n, K, H, W, seed, offset):
    """Instances carrying raw mask-head logits [n, K + offset, 28, 28] (blob-shaped, so that the
    thresholded masks are realistic) next to the probabilities mask_rcnn_inference derives."""
    base = synth.blob_instances(0, n, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed)
    p = base.pred_masks[:, 0].clamp(1e-6, 1 - 1e-6)
    own = torch.log(p) - torch.log1p(-p)                         # logit of the blob
    logits = torch.randn((n, K + offset, 28, 28), generator=g) * 3.0
    cls = base.pred_classes % K
    logits[torch.arange(n), cls + offset] = own
    # saturated / special values must survive the fused sigmoid exactly as torch's
    logits[0, cls[0] + offset, 0, :6] = torch.tensor([0.0, 88.0, -88.0, 104.0, -104.0, 1e-8])
    inst = uwcv.Instances((H, W), pred_boxes=base.pred_boxes, scores=base.scores, pred_classes=cls)
    inst.set("pred_mask_logits", logits)
    return inst


def test_fused_sigmoid_class_select_equals_mask_rcnn_inference():
    """uwcv_paste_measure_heads (channel select + sigmoid inside the paste kernel) == Detectron2's
    mask_rcnn_inference followed by the plain path, bit for bit, with torch's CUDA sigmoid (the
    device the reference's model runs on); against the CPU sigmoid the masks may differ where a
    1-ulp difference of a probability crosses the threshold -- counted and bounded."""
    H = W = 320
    for K, offset in ((4, 0), (4, 1), (1, 0)):
        batch = [_logit_instances(40 + 5 * k, K, H, W, seed=300 + k, offset=offset) for k in range(3)]
        fused = uwcv.measure_instances(batch, (H, W), mask_channel_offset=offset, return_planes=True)
        # unfused on the GPU: torch ops do what mask_rcnn_inference does
        plain = []
        for b in batch:
            lg = b.pred_mask_logits.cuda()
            idx = torch.arange(len(b), device="cuda")
            ch = (b.pred_classes.cuda() + offset) if lg.shape[1] > 1 else torch.zeros_like(idx)
            probs = lg[idx, ch][:, None].sigmoid().cpu()
            plain.append(uwcv.Instances((H, W), pred_boxes=b.pred_boxes, scores=b.scores,
                                        pred_classes=b.pred_classes, pred_masks=probs))
        want = uwcv.measure_instances(plain, (H, W), return_planes=True)
        assert np.array_equal(fused[0].ints, want[0].ints)
        assert np.array_equal(fused[0].floats, want[0].floats, equal_nan=True)
        assert torch.equal(fused[1], want[1])
        # oracle (CPU): mask_rcnn_inference restatement + the CPU pipeline
        if offset == 0:
            ob = [P.to_oracle_instances(uwcv.Instances((H, W), pred_boxes=b.pred_boxes, scores=b.scores,
                                                       pred_classes=b.pred_classes,
                                                       pred_masks=torch.zeros(len(b), 1, 28, 28)))
                  for b in batch]
            d2.mask_rcnn_inference(torch.cat([b.pred_mask_logits for b in batch]), ob)
            cpu_batch = [uwcv.Instances((H, W), pred_boxes=b.pred_boxes, scores=b.scores,
                                        pred_classes=b.pred_classes, pred_masks=o.pred_masks)
                         for b, o in zip(batch, ob)]
            ri, rf = P.oracle_table(cpu_batch, (H, W))
            diff = int((fused[0].ints[:, IC["area_px"]] != ri[:, IC["area_px"]]).sum())
            print(f"K={K}: instances whose pixel area differs from the CPU-sigmoid oracle: {diff} / {len(ri)}")
            assert diff <= max(1, len(ri) // 100)
            if diff == 0:
                compare_tables(fused[0], ri, rf)


def test_bad_class_channel_gives_empty_mask():
    inst = _logit_instances(10, 4, 128, 128, seed=5, offset=0)
    inst.pred_classes[3] = 17                                    # outside the head's channels
    t = uwcv.measure_instances(inst, (128, 128))
    assert t["valid"][3] == 0 and t["area_px"][3] == 0
    assert (t["valid"][[0, 1, 2]] == 1).all()


def test_single_forward_torchvision_maskrcnn():
    """SingleForward on a random-init torchvision Mask R-CNN: the batched NMS equals the
    Detectron2 restatement on the same head outputs, and the measured table equals the one
    obtained by materialising mask_rcnn_inference's probabilities first."""
    import torchvision
    torch.manual_seed(0)
    model = torchvision.models.detection.maskrcnn_resnet50_fpn(
        weights=None, weights_backbone=None, num_classes=5, min_size=256, max_size=256,
        box_score_thresh=0.0, rpn_post_nms_top_n_test=200).cuda().eval()
    images = [torch.rand(3, 256, 256) for _ in range(2)]
    sf = uwcv.SingleForward(model, score_thresh=0.05, nms_thresh=0.5, detections_per_image=50)
    inst = sf.predict(images)
    assert len(inst) == 2 and all(i.has("pred_mask_logits") for i in inst)
    assert sum(len(i) for i in inst) > 0
    table = sf.measure(images)
    assert 0 < len(table) <= sum(len(i) for i in inst)      # detector_postprocess drops empty boxes
    plain = []
    for i in inst:
        lg = i.pred_mask_logits
        probs = lg[torch.arange(len(i), device=lg.device), i.pred_classes + 1][:, None].sigmoid()
        plain.append(uwcv.Instances(i.image_size, pred_boxes=i.pred_boxes, scores=i.scores,
                                    pred_classes=i.pred_classes, pred_masks=probs))
    want = uwcv.measure_instances(plain, (256, 256))
    assert np.array_equal(table.ints, want.ints)
    assert np.array_equal(table.floats, want.floats, equal_nan=True)
    # the batched NMS against the oracle's fast_rcnn_inference on random head outputs
    g = torch.Generator().manual_seed(1)
    boxes_l, scores_l, shapes = [], [], []
    for b in range(3):
        R, K = 150 + 20 * b, 4
        ctr = torch.rand((R, 1, 2), generator=g) * 200 + 20
        wh = torch.rand((R, K, 2), generator=g) * 60 + 4
        bx = torch.cat((ctr - wh / 2, ctr + wh / 2), dim=2).reshape(R, K * 4)
        sc = torch.softmax(torch.randn((R, K + 1), generator=g) * 2, dim=1)
        boxes_l.append(bx); scores_l.append(sc); shapes.append((240, 250))
    got, got_rows = uwcv.fast_rcnn_inference(boxes_l, scores_l, shapes, 0.3, 0.5, 40)
    ref, ref_rows = d2.fast_rcnn_inference(boxes_l, scores_l, shapes, 0.3, 0.5, 40)
    for a, ar, r, rr in zip(got, got_rows, ref, ref_rows):
        assert torch.equal(a.pred_boxes.tensor.cpu(), r.pred_boxes.tensor)
        assert torch.equal(a.scores.cpu(), r.scores)
        assert torch.equal(a.pred_classes.cpu(), r.pred_classes)
        assert torch.equal(ar.cpu(), rr)


# ---------------------------------------------------------------- clean-up + RLE (f2) ---
def _oracle_export(batch, names, H, W):
    from oracle import cleanup as OC
    masks_l, scores_l = [], []
    for b in batch:
        res = d2.detector_postprocess(P.to_oracle_instances(b), H, W, 0.5)
        masks_l.append(res.pred_masks.numpy())
        scores_l.append(res.scores.numpy())
    return OC.export_rows(names, masks_l, scores_l, (H, W))


def _sorted_by_score(inst):
    order = torch.argsort(inst.scores, descending=True)
    out = uwcv.Instances(inst.image_size)
    for k, v in inst.get_fields().items():
        out.set(k, uwcv.Boxes(v.tensor[order]) if hasattr(v, "tensor") else v[order])
    return out


def test_clean_masks_and_rle_equal_the_reference_export():
    """export_rle == postprocess_masks + rle_encoding of the reference (oracle/cleanup.py) on the
    masks the Detectron2 restatement pastes: identical ImageId / EncodedPixels rows."""
    H, W = 192, 224                                        # W a multiple of 32, H not
    batch, names = [], []
    for k in range(4):                                     # dense: overlaps, cuts, several pieces
        batch.append(_sorted_by_score(synth.blob_instances(k, 60, H, W, seed=700 + k)))
        names.append(f"img{k}.tif")
    # hand-made image: border-touching boxes, near-border growth, a full-height all-ones mask
    # (runs continue across columns), identical boxes (the second is cut away entirely)
    boxes = torch.tensor([[10., 0., 42., float(H)], [0., 0., 30.5, 20.], [1., 150., 40., 191.],
                          [100., 60., 160., 120.], [100., 60., 160., 120.], [180., 100., float(W), 140.],
                          [120., 1., 150., 30.]])
    m = torch.ones((7, 1, 28, 28))
    g = torch.Generator().manual_seed(9)
    m[3:5] = synth.blob_probs(2, g)[:, None]
    m[6, 0, 10:18, 10:18] = 0.0                            # a hole: filled
    hand = uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(boxes),
                          scores=torch.linspace(0.95, 0.6, 7), pred_classes=torch.zeros(7, dtype=torch.int64),
                          pred_masks=m)
    batch.append(hand); names.append("hand.tif")
    zero = _sorted_by_score(synth.blob_instances(9, 12, H, W, seed=710))
    zero.scores[-1] = 0.0                                  # ``ori_score.all() < 0.5``: image skipped
    batch.append(zero); names.append("zero.tif")
    thin = uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(torch.tensor([[50., 20. + 9 * i, 52., 28. + 9 * i] for i in range(5)])),
                          scores=torch.linspace(0.9, 0.5, 5), pred_classes=torch.zeros(5, dtype=torch.int64),
                          pred_masks=torch.ones((5, 1, 28, 28)))
    batch.append(thin); names.append("thin.tif")           # 2 occupied columns < 5 instances: truncated
    batch.append(uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(torch.zeros((0, 4))), scores=torch.zeros(0),
                                pred_classes=torch.zeros(0, dtype=torch.int64),
                                pred_masks=torch.zeros((0, 1, 28, 28))))
    names.append("none.tif")
    ids, enc = _oracle_export(batch, names, H, W)
    got = uwcv.export_rle(batch, (H, W), names)
    assert got.image_id == ids
    bad = [k for k, (a, b) in enumerate(zip(got.encoded_pixels, enc)) if a != b]
    assert not bad, f"{len(bad)} of {len(enc)} rows differ, first {bad[:5]} ({got.image_id[bad[0]]})"
    assert "zero" not in ids and ids.count("thin") == 2 and "none" not in ids
    n_empty = sum(1 for e in enc if e == "")
    n_multi = int(got.multi_piece.sum())
    print(f"{len(enc)} rows, {n_empty} emptied ({n_multi} as multi-piece)")
    assert n_multi > 0 and n_empty > n_multi               # both mechanisms are exercised
    # areas reported by the kernel == pixels of the decoded strings
    from oracle import cleanup as OC
    for k in (0, len(enc) // 2, len(enc) - 1):
        assert int(OC.rle_decode(enc[k], (W, H)).sum()) == int(got.area[k])


def test_rle_full_size_round_trip():
    """Config-2-sized image: every exported run lies inside its instance's pixel box, the decoded
    masks of an image are pairwise disjoint (the point of the overlap cut), areas match."""
    from oracle import cleanup as OC
    H = W = 2048
    inst = _sorted_by_score(synth.blob_instances(0, 1000, H, W, seed=1234))
    got = uwcv.export_rle(inst, (H, W), ["a.tif"])
    assert len(got) == 1000
    cover = np.zeros(H * W, dtype=np.uint8)
    for k in range(len(got)):
        s = np.array(got.encoded_pixels[k].split(), dtype=np.int64).reshape(-1, 2)
        assert int(s[:, 1].sum()) == int(got.area[k])
        for a, l in s:
            cover[a - 1:a - 1 + l] += 1
    assert cover.max() <= 1


def test_tile_major_shards_gather_into_the_whole_image_table():
    """configs[3] partition of one micrograph: measuring the tile-major shards (as 4 ranks would)
    and sorting the concatenated rows gives the table of the unsharded call."""
    H = W = 1024
    inst = synth.blob_instances(0, 600, H, W, seed=42)
    whole = uwcv.measure_instances(inst, (H, W))
    parts = []
    boxes = api.scale_clip_boxes(inst.pred_boxes.tensor, inst.image_size, (H, W))[0]
    for r in range(4):
        idx = uwcv.shard_instances_by_tile(boxes, (H, W), r, 4)
        parts.append(uwcv.measure_instances(uwcv.take_instances(inst, idx), (H, W)))
    assert sum(len(p) for p in parts) == len(whole) and min(len(p) for p in parts) > 0
    cat = uwcv.MeasurementTable.concat(parts)
    gi, gf = uwcv.sort_rows(torch.from_numpy(cat.ints), torch.from_numpy(cat.floats))
    assert np.array_equal(gi.numpy(), whole.ints)
    assert np.array_equal(gf.numpy(), whole.floats, equal_nan=True)


def test_clean_masks_on_arbitrary_binary_shapes():
    """Random 0/1 patterns pasted one-to-one (28 x 28 boxes on integer coordinates reproduce the
    pattern pixel for pixel): holes of every connectivity, diagonal bridges, spurs, masks on the
    image border, heavy overlaps.  Rows must equal the reference's postprocess_masks + rle_encoding,
    and the per-instance measurement rows of the same shapes must equal the oracle's."""
    H, W = 96, 128
    g = torch.Generator().manual_seed(2024)
    batch, names = [], []
    for k, density in enumerate((0.35, 0.5, 0.62, 0.75, 0.9)):
        n = 24
        pat = (torch.rand((n, 28, 28), generator=g) < density).float()
        # a few solid frames with diagonal pinholes (4-connected fill vs 8-connected pieces)
        pat[0] = 1.0
        pat[0, 5:9, 5:9] = 0.0
        pat[0, 9, 9] = 0.0                                 # hole touching another hole diagonally
        x0 = torch.randint(-6, W - 20, (n,), generator=g).float()
        y0 = torch.randint(-6, H - 20, (n,), generator=g).float()
        boxes = torch.stack([x0, y0, x0 + 28, y0 + 28], dim=1)
        inst = uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(boxes),
                              scores=torch.linspace(0.99, 0.51, n),
                              pred_classes=torch.randint(0, 4, (n,), generator=g),
                              pred_masks=pat[:, None])
        batch.append(inst); names.append(f"p{k}.tif")
    ids, enc = _oracle_export(batch, names, H, W)
    got = uwcv.export_rle(batch, (H, W), names)
    assert got.image_id == ids
    bad = [k for k, (a, b) in enumerate(zip(got.encoded_pixels, enc)) if a != b]
    assert not bad, f"{len(bad)} of {len(enc)} rows differ, first {bad[:5]}"
    assert 0 < int(got.multi_piece.sum()) < len(enc)
    # the same shapes through the measurement path (contour tracing of ragged shapes)
    table = uwcv.measure_instances(batch, (H, W))
    ri, rf = P.oracle_table(batch, (H, W))
    compare_tables(table, ri, rf)


def test_single_forward_stream_equals_per_batch_calls():
    """SingleForward.measure_stream (network of batch i + 1 enqueued while batch i is measured)
    yields the tables of measure() batch by batch."""
    import torchvision
    torch.manual_seed(1)
    model = torchvision.models.detection.maskrcnn_resnet50_fpn(
        weights=None, weights_backbone=None, num_classes=5, min_size=192, max_size=192,
        box_score_thresh=0.0, rpn_post_nms_top_n_test=100).cuda().eval()
    sf = uwcv.SingleForward(model, score_thresh=0.05, nms_thresh=0.5, detections_per_image=40)
    g = torch.Generator().manual_seed(3)
    batches = [[torch.rand((3, 192, 192), generator=g) for _ in range(2)] for _ in range(4)]
    want = [sf.measure(b) for b in batches]
    got = list(sf.measure_stream(batches))
    assert len(got) == len(want) and sum(len(t) for t in want) > 0
    for a, b in zip(got, want):
        assert np.array_equal(a.ints, b.ints)
        assert np.array_equal(a.floats, b.floats, equal_nan=True)


def test_postprocess_masks_literal_drop_in():
    """uwcv.postprocess_masks(ori_mask, ori_score, image) -- N x H x W bool masks in, list of
    cleaned uint8 masks out -- equals the reference's function (oracle/cleanup.py) mask for mask,
    including its None / [] / truncation behaviour."""
    from oracle import cleanup as OC
    rng = np.random.default_rng(11)
    for H, W in ((96, 128), (77, 131)):                       # W a multiple of 16 (vector path) and not
        n = 30
        masks = np.zeros((n, H, W), dtype=bool)
        for i in range(n):
            y0, x0 = rng.integers(0, H - 12), rng.integers(0, W - 12)
            h, w = rng.integers(6, 40), rng.integers(6, 40)
            blob = rng.random((min(h, H - y0), min(w, W - x0))) < 0.8
            masks[i, y0:y0 + blob.shape[0], x0:x0 + blob.shape[1]] = blob
        masks[3] = False                                      # an empty mask stays in the list
        masks[5, :, 10:14] = True                             # full height: touches both borders
        scores = np.linspace(0.95, 0.55, n)
        want = OC.postprocess_masks(masks.copy(), scores, (H, W))
        got = uwcv.postprocess_masks(masks, scores, np.zeros((H, W, 3), np.uint8))
        assert len(got) == len(want) == n
        for k, (a, b) in enumerate(zip(got, want)):
            assert a.dtype == np.uint8 and np.array_equal(a, b), (H, W, k)
        assert sum(1 for b in want if b.sum() == 0) > 3        # overlaps / pieces were exercised
        # torch input, zero score -> None; no masks -> None; all-empty masks -> []
        assert uwcv.postprocess_masks(torch.from_numpy(masks), np.append(scores[:-1], 0.0), (H, W)) is None
        assert uwcv.postprocess_masks(masks[:0], scores[:0], (H, W)) is None
        assert uwcv.postprocess_masks(np.zeros((3, H, W), bool), scores[:3], (H, W)) == []
        # fewer occupied columns than instances: the list is truncated (:277-284)
        thin = np.zeros((5, H, W), bool)
        thin[:, 20:40, 30:32] = True
        g2, w2 = uwcv.postprocess_masks(thin, scores[:5], (H, W)), OC.postprocess_masks(thin.copy(), scores[:5], (H, W))
        assert len(g2) == len(w2) == 2 and all(np.array_equal(a, b) for a, b in zip(g2, w2))


def test_fused_gather_branch_on_one_gpu():
    """The GatherDst branch of the trace kernel (contour.cu: rows stored into every peer's table at
    this rank's row offset) exercised on ONE device: three 'ranks' are three calls on different
    inputs, the three 'peer' tables are three local buffers; every table must end up holding the
    concatenation of the three plain results."""
    from uwcv import _lib
    dev = torch.device("cuda", 0)
    eng = uwcv.Engine.get(dev)
    H, W = 300, 420
    world = 3
    parts = []
    for r in range(world):
        inst = synth.blob_instances(r, 37 + 11 * r, H, W, seed=90)
        b, keep = api.scale_clip_boxes(inst.pred_boxes.tensor, inst.image_size, (H, W))
        parts.append((inst.pred_masks[keep, 0].contiguous().to(dev), b[keep].contiguous().to(dev),
                      inst.scores[keep].to(dev), inst.pred_classes[keep].to(dev)))
    counts = [int(p[1].shape[0]) for p in parts]
    total = sum(counts)
    tabs_i = [torch.full((total, schema.NUM_INT), -7, dtype=torch.int64, device=dev) for _ in range(world)]
    tabs_f = [torch.full((total, schema.NUM_FLOAT), -7.0, dtype=torch.float64, device=dev) for _ in range(world)]
    plain = []
    base = 0
    for r, (m, b, s, c) in enumerate(parts):
        g = _lib.Gather()
        g.world, g.row_base = world, base
        for p in range(world):
            g.rows_i[p] = tabs_i[p].data_ptr()
            g.rows_f[p] = tabs_f[p].data_ptr()
        img = torch.full((counts[r],), r, dtype=torch.int32, device=dev)
        ri, rf, _ = eng.run(m, b, H, W, image_idx=img, classes=c, scores=s, gather=g)
        eng.check_status()
        plain.append((ri.clone(), rf.clone()))
        base += counts[r]
    torch.cuda.synchronize()
    want_i = torch.cat([p[0] for p in plain])
    want_f = torch.cat([p[1] for p in plain])
    for p in range(world):
        assert torch.equal(tabs_i[p], want_i), p
        assert torch.equal(tabs_f[p].nan_to_num(), want_f.nan_to_num()), p
    # and the rows equal a call without the gather
    ri, rf, _ = eng.run(parts[1][0], parts[1][1], H, W, classes=parts[1][3], scores=parts[1][2],
                        image_idx=torch.full((counts[1],), 1, dtype=torch.int32, device=dev))
    assert torch.equal(ri, plain[1][0]) and torch.equal(rf.nan_to_num(), plain[1][1].nan_to_num())


def test_device_resident_inputs_take_the_sync_free_path_and_fall_back():
    """Predictor output that already lives on the GPU (the single-forward path) is measured without
    a host synchronisation: 'no box is dropped' is assumed and verified by a device flag read back
    with the rows.  Same table as with host inputs; a batch with an empty box falls back to the
    general path and still equals the host result (the box is dropped, as Boxes.nonempty does)."""
    dev = torch.device("cuda", 0)
    H, W = 240, 336
    batch = [synth.blob_instances(k, 30 + 5 * k, H, W, seed=410) for k in range(3)]

    def to_dev(b):
        out = []
        for inst in b:
            o = uwcv.Instances(inst.image_size)
            for k, v in inst.get_fields().items():
                o.set(k, uwcv.Boxes(v.tensor.to(dev)) if hasattr(v, "tensor") else v.to(dev))
            out.append(o)
        return out

    want = uwcv.measure_instances(batch, (H, W), device=dev)
    got = uwcv.measure_instances(to_dev(batch), (H, W), device=dev)
    assert np.array_equal(got.ints, want.ints) and np.array_equal(got.floats, want.floats, equal_nan=True)
    stream = uwcv.MeasurementStream(dev, depth=2)
    for t in stream.map([to_dev(batch)] * 3, (H, W)):
        assert np.array_equal(t.ints, want.ints)
    # an empty (zero-width) box in image 1: dropped by detector_postprocess
    bad = [synth.blob_instances(k, 30 + 5 * k, H, W, seed=410) for k in range(3)]
    bad[1].pred_boxes.tensor[4, 2] = bad[1].pred_boxes.tensor[4, 0]
    want = uwcv.measure_instances(bad, (H, W), device=dev)
    got = uwcv.measure_instances(to_dev(bad), (H, W), device=dev)
    assert len(want) == sum(len(b) for b in batch) - 1
    assert np.array_equal(got.ints, want.ints) and np.array_equal(got.floats, want.floats, equal_nan=True)


def test_moment_overflow_is_refused_not_wrapped():
    """m30 = sum x^3 of a near-full-frame mask passes int64 above ~6 000 px a side; such a
    tile is refused with the device status UWCV_E_TOO_LARGE instead of wrapping silently, while an
    8192-pixel frame with ordinary boxes is still measured."""
    from uwcv import _lib
    dev = torch.device("cuda", 0)
    eng = uwcv.Engine.get(dev)
    H = W = 8192
    m = torch.ones((2, 28, 28), device=dev)
    ok_boxes = torch.tensor([[100., 100., 400., 300.], [7000., 7800., 8192., 8192.]], device=dev)
    ri, rf, st = eng.run(m, ok_boxes, H, W)
    eng.check_status()
    want = [int(d2.paste_one_cropped(m[k].cpu(), ok_boxes[k].cpu(), H, W, 0.5)[0].sum()) for k in range(2)]
    assert ri[:, IC["area_px"]].tolist() == want and min(want) > 50000
    huge = torch.tensor([[0., 0., 8192., 8192.], [10., 10., 50., 50.]], device=dev)
    eng.run(m, huge, H, W)
    with pytest.raises(_lib.UwcvError) as ei:
        eng.check_status()
    assert ei.value.code == -8
    eng.run(m, ok_boxes, H, W)                       # the engine is usable afterwards
    eng.check_status()


def _shared_table_worker(q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "uw-com-vision_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gc
    import torch.distributed as dist
    from uwcv.dist import SharedHostTable
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(29700 + os.getpid() % 1000)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        ht = SharedHostTable(dev)
        kept, ok = [], True
        for call in range(2 * ht.SETS + 1):                    # every set is reused at least once
            n = 50 + call
            ri = torch.arange(n * 20, dtype=torch.int64, device=dev).view(n, 20) + 1000 * call
            rf = (torch.arange(n * 30, dtype=torch.float64, device=dev).view(n, 30) + 0.5) * (call + 1)
            k, seq, base, total = ht.begin([n])
            assert (base, total) == (0, n) and seq == call and k == call % ht.SETS
            ht.copy_rows(k, seq, base, ri, rf)
            torch.cuda.current_stream().synchronize()
            ht.wait_all(k, seq)                                 # the flag word landed behind the rows
            a_i, a_f = ht.arrays(k, seq, 0, total, whole=True)
            ok &= bool(np.array_equal(a_i, ri.cpu().numpy()) and np.array_equal(a_f, rf.cpu().numpy()))
            kept.append((a_i[:, 3], a_f))                       # a slice keeps the set leased ...
            if len(kept) > ht.SETS - 2:
                kept.pop(0)                                     # ... until it is dropped
                gc.collect()
        # a table that is never dropped blocks the reuse of its set with a clear error
        kept.clear()
        del a_i, a_f
        gc.collect()
        hold = [ht.arrays(*(ht.begin([10])[:2]), 0, 10, whole=True) for _ in range(ht.SETS)]
        try:
            ht.begin([10], timeout_s=0.3)
            blocked = False
        except RuntimeError as e:
            blocked = "still referenced" in str(e)
        q.put((ok, blocked))
    finally:
        dist.destroy_process_group()


def test_shared_host_table_on_one_gpu():
    """uwcv.dist.SharedHostTable (the host-sink of the e2e gather) with a one-rank group: rows and
    the completion flag land in the registered shared segment, sets are reused only after their
    views are dropped, and a table held forever is reported instead of being overwritten."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_shared_table_worker, args=(q,))
    p.start()
    ok, blocked = q.get(timeout=180)
    p.join(timeout=60)
    assert p.exitcode == 0
    assert ok, "rows read back from the shared table differ"
    assert blocked, "reusing a set whose table is still referenced must raise"


def test_split_pipeline_equals_fused_kernel():
    """Stage 8 (planes written from the tiles by plane_fill_kernel, on its own stream) against the
    fused paste kernel of the single-stream call: identical planes and rows -- bands taller than the
    shared-memory band image (chunked), frame-sized and empty tiles, W not a multiple of 32,
    instance ranges, both workspaces of Engine.run_overlapped."""
    dev = torch.device("cuda", 0)
    eng = api.Engine.get(dev)
    g = torch.Generator().manual_seed(77)
    for (H, W, n) in ((300, 4000, 90), (144, 208, 300), (64, 33, 40)):
        cx = torch.rand(n, generator=g) * (W + 20) - 10
        cy = torch.rand(n, generator=g) * (H + 20) - 10
        w = torch.exp(torch.rand(n, generator=g) * 8 - 2).clamp(max=2.0 * W)
        h = torch.exp(torch.rand(n, generator=g) * 8 - 2).clamp(max=2.0 * H)
        boxes = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
        boxes[0] = torch.tensor([0., 0., float(W), float(H)])            # whole frame: the tallest band
        boxes[1] = torch.tensor([5., 5., 5., 40.])                       # zero width: empty tile
        boxes[2] = torch.tensor([W - 3.5, 0., float(W), float(H)])       # right edge, full height
        boxes[:, 0::2] = boxes[:, 0::2].clamp(0, W)
        boxes[:, 1::2] = boxes[:, 1::2].clamp(0, H)
        masks = torch.rand(n, 28, 28, generator=g)
        masks[0] = 1.0
        masks[::6] = (masks[::6] > 0.5).float()
        d_b, d_m = boxes.contiguous().to(dev), masks.contiguous().to(dev)
        words = api.tile_words(boxes, H, W)
        sc = torch.rand(n, generator=g).to(dev)

        def fresh():
            return (eng.alloc_planes(n, H, W).fill_(-1), torch.empty((n, 20), dtype=torch.int64, device=dev),
                    torch.empty((n, 30), dtype=torch.float64, device=dev))

        p0, i0, f0 = fresh()
        eng.run(d_m, d_b, H, W, planes=p0, rows_i=i0, rows_f=f0, scores=sc, n_tile_words=words)   # fused, stages = 7
        torch.cuda.synchronize()
        assert int(eng.status.cpu()[0]) == 0
        # (a) the split stages on one stream
        p1, i1, f1 = fresh()
        eng.run(d_m, d_b, H, W, planes=p1, rows_i=i1, rows_f=f1, scores=sc, n_tile_words=words,
                stages=1 | 2 | 16 | 8 | 4)
        torch.cuda.synchronize()
        assert torch.equal(p1, p0) and torch.equal(i1, i0) and torch.equal(f1.nan_to_num(), f0.nan_to_num())
        # (b) three streams, instance ranges, both workspaces in turn
        third = n // 3
        ranges = [(0, third, None), (third, third, None), (2 * third, n - 2 * third, None)]
        for rep in range(2):
            p2, i2, f2 = fresh()
            eng.run_overlapped(d_m, d_b, H, W, planes=p2, rows_i=i2, rows_f=f2, scores=sc, n_tile_words=words,
                               paste_ranges=ranges if rep else None, split=True)
            assert eng.planes_done is not None
            torch.cuda.synchronize()
            assert torch.equal(p2, p0), (H, W, rep)
            assert torch.equal(i2, i0) and torch.equal(f2.nan_to_num(), f0.nan_to_num())
        # the planes are the Detectron2-literal masks
        ref = d2.paste_masks_in_image(masks, boxes, (H, W))
        assert torch.equal(eng.unpack(p2, H, W).cpu(), ref)


def test_split_pipeline_streams_different_batches():
    """Calls in flight on the three streams of the split pipeline (tile kernel of call i + 1, plane
    fill and border trace of call i, two workspaces in turn) with a DIFFERENT batch per call: every
    call returns its own planes and rows (a fill or a trace that read a workspace the next layout had
    already handed out again would show here)."""
    H, W = 384, 512
    batches = [[synth.blob_instances(10 * s + k, 150 + 40 * ((s + k) % 4), H, W, seed=900 + 7 * s + k,
                                     size_range=(6.0, 120.0)) for k in range(3)] for s in range(7)]
    want = []
    for b in batches:                                   # reference: the fused single-stream kernels
        masks = torch.cat([i.pred_masks[:, 0] for i in b])
        boxes = torch.cat([api.scale_clip_boxes(i.pred_boxes.tensor, (H, W), (H, W))[0] for i in b])
        want.append(plane_crcs(uwcv.paste_masks_in_image(masks, boxes, (H, W), packed=True)))
    rows = [uwcv.measure_instances(b, (H, W)) for b in batches]          # rows-only contract
    # (three images per call = three chunks of tiles; one plane fill per chunk or one per call)
    for depth, fills in ((2, None), (3, None), (2, "once"), (3, "per_chunk")):
        stream = uwcv.MeasurementStream(depth=depth, fills=fills)
        assert stream.fill_once == ((depth >= 3) if fills is None else fills == "once")
        got = list(stream.map(batches, (H, W), return_planes=True))
        for (table, planes), crc, ref in zip(got, want, rows):
            assert len(table) == len(crc)
            assert np.array_equal(plane_crcs(planes), crc)
            assert np.array_equal(table.ints, ref.ints)
            assert np.array_equal(table.floats, ref.floats, equal_nan=True)


def test_compressible_plane_buffers():
    """uwcv_planes_alloc: planes written into compressible device memory are the planes written into
    ordinary memory, bit for bit, for the fused kernel and for the split pipeline; the allocation is
    released with its last tensor; the raw calls round-trip."""
    import ctypes as C
    from uwcv import _lib
    dev = torch.device("cuda", 0)
    eng = api.Engine.get(dev)
    L = _lib.lib()
    ptr, comp = C.c_void_p(), C.c_int(-1)
    assert L.uwcv_planes_alloc(3 << 20, C.byref(ptr), C.byref(comp)) == 0 and ptr.value and comp.value in (0, 1)
    assert ptr.value % (2 << 20) == 0
    assert L.uwcv_planes_free(ptr) == 0 and L.uwcv_planes_free(ptr) == -2   # second free: unknown pointer
    H, W, n = 300, 500, 120
    inst = synth.blob_instances(3, n, H, W, seed=41, size_range=(4.0, 200.0))
    masks, boxes = inst.pred_masks[:, 0].contiguous().to(dev), inst.pred_boxes.tensor.contiguous().to(dev)
    wpr = L.uwcv_plane_row_words(W)
    plain = torch.empty((n, H, wpr), dtype=torch.int32, device=dev).fill_(-1)
    eng.run(masks, boxes, H, W, planes=plain)
    for split in (False, True):
        pc = eng.alloc_planes(n, H, W)
        assert pc.is_cuda and pc.shape == plain.shape and pc.dtype == torch.int32
        pc.fill_(-1)
        if split:
            eng.run_overlapped(masks, boxes, H, W, planes=pc, split=True)
        else:
            eng.run(masks, boxes, H, W, planes=pc)
        torch.cuda.synchronize()
        assert torch.equal(pc, plain)
        assert torch.equal(eng.unpack(pc, H, W).cpu(), d2.paste_masks_in_image(inst.pred_masks[:, 0], inst.pred_boxes.tensor, (H, W)))
        view = pc[5:9]                               # a view keeps the allocation alive
        del pc
        assert torch.equal(view, plain[5:9])
        del view
    if eng.compressible_planes:                      # granted on this device: the buffers above were compressible
        buf = api._CompressibleBuffer.create(eng, 1 << 18)
        assert buf is not None and buf.compressed
        del buf
