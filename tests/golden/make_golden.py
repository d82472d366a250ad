"""Generate the golden fixtures under tests/golden/ (run in the build container; the GPU
box only reads the committed files).

The reference ships no tests, fixtures or golden vectors, and its
scripts cannot be imported.  The fixtures below are outputs of the third-party libraries
the reference calls, as installed in this image -- torch 2.11 CPU ``grid_sample`` through
the restated Detectron2 glue, OpenCV 4.13, torchvision 0.26 NMS -- on seeded inputs:

  c1_maskrcnn.npz   config 1: random-init torchvision Mask R-CNN R50-FPN on a 1024x1024
                    grayscale image, 200 detections (raw head output) + oracle rows +
                    per-instance CRC of the packed full-frame plane
  blobs_small.npz   seeded blob predictions (3 images, 320x416 output, rescaled from a
                    256x333 network input) + oracle rows + plane CRCs
  nms_small.npz     clustered candidates + torchvision _batched_nms_vanilla keep list
  kat.json          hand-checkable known answers of SURVEY.md section 8(c), re-derived here
                    with cv2 (the test asserts both the stored numbers and the survey's)

Usage:  python tests/golden/make_golden.py [--skip-c1]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))

import cv2  # noqa: E402
from oracle import d2, measure as M, pipeline as P  # noqa: E402
from uwcv import synth  # noqa: E402
from uwcv.structures import Boxes, Instances  # noqa: E402


def plane_crcs(res: d2.Instances, thr=0.5):
    """CRC32 of every instance's packed full-frame bit-plane (row stride as the product)."""
    H, W = res.image_size
    out = []
    for win, y0, x0 in P.oracle_windows(res, thr):
        full = np.zeros((H, W), dtype=bool)
        full[y0:y0 + win.shape[0], x0:x0 + win.shape[1]] = win
        out.append(zlib.crc32(P.pack_bits(full).tobytes()))
    return np.array(out, dtype=np.int64)


def make_c1():
    """SURVEY.md 8(d) C1.  labels 1..4 -> classes 0..3."""
    import torchvision
    d2.assert_cpu_capability()
    torch.manual_seed(0)
    img = torch.rand(1, 1024, 1024).expand(3, -1, -1)
    model = torchvision.models.detection.maskrcnn_resnet50_fpn(
        weights=None, weights_backbone=None, num_classes=5, box_score_thresh=0.0,
        box_detections_per_img=200, min_size=1024, max_size=1024).eval()
    captured = {}
    rh = torchvision.models.detection.roi_heads
    orig = rh.maskrcnn_inference

    def hook(x, labels):
        out = orig(x, labels)
        captured["probs"] = [o.detach().clone() for o in out]
        return out

    rh.maskrcnn_inference = hook
    try:
        with torch.no_grad():
            det = model([img])[0]
    finally:
        rh.maskrcnn_inference = orig
    probs = captured["probs"][0]                      # (200, 1, 28, 28)
    boxes = det["boxes"].detach().float()
    scores = det["scores"].detach().float()
    classes = (det["labels"].detach() - 1).to(torch.int64)
    inst = Instances((1024, 1024))
    inst.pred_boxes = Boxes(boxes)
    inst.scores = scores
    inst.pred_classes = classes
    inst.pred_masks = probs.float()
    ri, rf = P.oracle_table([inst], (1024, 1024))
    res = P.postprocess_boxes(P.to_oracle_instances(inst), (1024, 1024))
    crcs = plane_crcs(res)
    np.savez_compressed(os.path.join(HERE, "c1_maskrcnn.npz"),
                        masks=probs.numpy().astype(np.float32), boxes=boxes.numpy(),
                        scores=scores.numpy(), classes=classes.numpy(),
                        rows_i=ri, rows_f=rf, plane_crc=crcs)
    print("c1:", len(inst), "instances ->", ri.shape[0], "rows; empty masks:",
          int((ri[:, 3] == 0).sum()))


def make_blobs():
    H, W = 320, 416
    batch = []
    for k in range(3):
        inst = synth.blob_instances(k, 40, 256, 333, seed=77, size_range=(4.0, 90.0))
        batch.append(inst)
    ri, rf = P.oracle_table(batch, (H, W))
    crcs = []
    for inst in batch:
        res = P.postprocess_boxes(P.to_oracle_instances(inst), (H, W))
        crcs.append(plane_crcs(res))
    np.savez_compressed(os.path.join(HERE, "blobs_small.npz"), rows_i=ri, rows_f=rf,
                        plane_crc=np.concatenate(crcs), H=H, W=W)
    print("blobs:", ri.shape[0], "rows")


def make_nms():
    b, s, c = synth.clustered_candidates(600, 512, 512, seed=5, n_clusters=8)
    keep = d2.batched_nms_vanilla(b, s, c, 0.5)
    sel = s[keep] > 0.05
    np.savez_compressed(os.path.join(HERE, "nms_small.npz"), boxes=b.numpy(), scores=s.numpy(),
                        classes=c.numpy(), keep_thr005=keep[sel].numpy(),
                        keep_all=keep.numpy())
    print("nms:", len(b), "candidates ->", int(sel.sum()), "kept above 0.05")


def make_kat():
    kat = {}

    def contour_facts(mask):
        cs, _ = cv2.findContours(mask.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        out = []
        for c in cs:
            r = cv2.minAreaRect(c)
            out.append(dict(points=c.reshape(-1, 2).tolist(), area=cv2.contourArea(c),
                            arclen=cv2.arcLength(c, True),
                            rect=[r[0][0], r[0][1], r[1][0], r[1][1], r[2]]))
        return out

    m = np.zeros((30, 40), np.uint8); m[10:20, 5:30] = 1
    kat["rect_10x25"] = contour_facts(m)
    yy, xx = np.mgrid[:200, :200]
    disc = (((xx - 100) ** 2 + (yy - 90) ** 2) <= 40 ** 2).astype(np.uint8)
    kat["disc_r40"] = dict(pixels=int(disc.sum()), contours=contour_facts(disc))
    m = np.zeros((8, 8), np.uint8); m[3, 3] = 1
    kat["single_pixel"] = contour_facts(m)
    m = np.zeros((8, 10), np.uint8); m[3, 2:8] = 1
    kat["line_1x6"] = contour_facts(m)
    m = np.zeros((9, 11), np.uint8); m[2:6, 2:6] = 1; m[3, 6:9] = 1
    kat["square_spur"] = contour_facts(m)
    m = np.zeros((12, 12), np.uint8); m[2:10, 2:10] = 1; m[4:8, 4:8] = 0
    kat["ring"] = contour_facts(m)
    m = np.zeros((8, 8), np.uint8); m[1:3, 1:3] = 1; m[3:5, 3:5] = 1
    kat["diag_squares"] = contour_facts(m)
    # reference-literal union KAT: three ellipses, union painted, rows left -> right
    masks = np.stack([((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1
                      for cx, cy, a, b in ((100, 90, 40, 40), (40, 40, 25, 12), (160, 160, 10, 30))])
    rows = M.get_mask_contours((200, 200, 3), np.array([0, 0, 0]), masks, [0])
    kat["union_three_ellipses"] = rows.tolist()
    # all-0.5 / all-ones paste KATs
    ones = torch.ones(1, 28, 28)
    half = torch.full((1, 28, 28), 0.5)
    kat["paste_all_ones_count"] = int(d2.paste_masks_in_image(
        ones, torch.tensor([[10., 20., 50., 60.]]), (100, 100)).sum())
    kat["paste_all_half_count"] = int(d2.paste_masks_in_image(
        half, torch.tensor([[10.5, 20.5, 50.5, 60.5]]), (100, 100)).sum())
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("kat: written")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-c1", action="store_true")
    a = ap.parse_args()
    print("torch", torch.__version__, "cv2", cv2.__version__, "cpu capability",
          torch.backends.cpu.get_cpu_capability())
    make_kat()
    make_nms()
    make_blobs()
    if not a.skip_c1:
        make_c1()
