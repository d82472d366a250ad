"""Seeded inputs of the reference-execution goldens (tests/golden/ref_exec_*.npz).

One place builds the inputs; ``tests/golden/make_ref_golden.py`` (build container: runs the
reference's own functions on them, oracle/ref_exec.py), ``tests/test_ref_exec.py`` (CPU: oracle ==
golden) and ``tests/test_gpu_ref_parity.py`` (GPU: CUDA path == golden) all import it, and every
golden file carries a CRC of the input bytes so that a drift of the generators shows up as
"inputs changed", not as a parity failure.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Tuple

import numpy as np
import torch

import uwcv
from uwcv import synth


def digest(instances) -> int:
    """CRC32 over the predictor outputs of a list of Instances."""
    c = 0
    for inst in instances:
        for t in (inst.pred_boxes.tensor, inst.scores, inst.pred_classes, inst.pred_masks):
            c = zlib.crc32(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes(), c)
    return c


def sorted_by_score(inst):
    order = torch.argsort(inst.scores, descending=True)
    out = uwcv.Instances(inst.image_size)
    for k, v in inst.get_fields().items():
        out.set(k, uwcv.Boxes(v.tensor[order]) if hasattr(v, "tensor") else v[order])
    return out


# --- measurement fixtures ----------------------------------------------------------------

def union_dense() -> Tuple[List, Tuple[int, int]]:
    """3 images x 120 overlapping / touching blobs at 384 x 512 (the f1 GPU test's input)."""
    H, W = 384, 512
    return [synth.blob_instances(k, 120, H, W, seed=300, size_range=(10.0, 110.0))
            for k in range(3)], (H, W)


def blobs_rescaled() -> Tuple[List, Tuple[int, int]]:
    """3 images x 40 blobs predicted at 256 x 333 and post-processed to 320 x 416
    (``detector_postprocess`` scales the boxes)."""
    return [synth.blob_instances(k, 40, 256, 333, seed=77, size_range=(4.0, 90.0))
            for k in range(3)], (320, 416)


def c1_maskrcnn(golden_dir: str) -> Tuple[List, Tuple[int, int]]:
    """configs[0]: raw head output of the random-init Mask R-CNN (stored by make_golden.py)."""
    import os
    g = np.load(os.path.join(golden_dir, "c1_maskrcnn.npz"))
    inst = uwcv.Instances((1024, 1024), pred_boxes=uwcv.Boxes(torch.from_numpy(g["boxes"])),
                          scores=torch.from_numpy(g["scores"]),
                          pred_classes=torch.from_numpy(g["classes"]),
                          pred_masks=torch.from_numpy(g["masks"]))
    return [inst], (1024, 1024)


def three_ellipses() -> Tuple[np.ndarray, np.ndarray]:
    """bool masks given directly (no paste), all class 0."""
    yy, xx = np.mgrid[:200, :200]
    masks = np.stack([((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1
                      for cx, cy, a, b in ((100, 90, 40, 40), (40, 40, 25, 12), (160, 160, 10, 30))])
    return masks, np.zeros(3, dtype=np.int64)


# --- clean-up / export fixtures -------------------------------------------------------------

def export_batch() -> Tuple[List, List[str], Tuple[int, int]]:
    """The f2 GPU test's folder of images: dense overlaps, border cases, the zero-score and
    column-count quirks, an image without detections."""
    H, W = 192, 224
    batch, names = [], []
    for k in range(4):
        batch.append(sorted_by_score(synth.blob_instances(k, 60, H, W, seed=700 + k)))
        names.append(f"img{k}.tif")
    boxes = torch.tensor([[10., 0., 42., float(H)], [0., 0., 30.5, 20.], [1., 150., 40., 191.],
                          [100., 60., 160., 120.], [100., 60., 160., 120.], [180., 100., float(W), 140.],
                          [120., 1., 150., 30.]])
    m = torch.ones((7, 1, 28, 28))
    g = torch.Generator().manual_seed(9)
    m[3:5] = synth.blob_probs(2, g)[:, None]
    m[6, 0, 10:18, 10:18] = 0.0
    batch.append(uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(boxes), scores=torch.linspace(0.95, 0.6, 7),
                                pred_classes=torch.zeros(7, dtype=torch.int64), pred_masks=m))
    names.append("hand.tif")
    zero = sorted_by_score(synth.blob_instances(9, 12, H, W, seed=710))
    zero.scores[-1] = 0.0
    batch.append(zero); names.append("zero.tif")
    thin = uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(torch.tensor([[50., 20. + 9 * i, 52., 28. + 9 * i] for i in range(5)])),
                          scores=torch.linspace(0.9, 0.5, 5), pred_classes=torch.zeros(5, dtype=torch.int64),
                          pred_masks=torch.ones((5, 1, 28, 28)))
    batch.append(thin); names.append("thin.tif")
    batch.append(uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(torch.zeros((0, 4))), scores=torch.zeros(0),
                                pred_classes=torch.zeros(0, dtype=torch.int64),
                                pred_masks=torch.zeros((0, 1, 28, 28))))
    names.append("none.tif")
    return batch, names, (H, W)


def bool_mask_cases() -> List[Dict]:
    """``postprocess_masks(ori_mask, ori_score, image)`` inputs: N x H x W bool + scores."""
    rng = np.random.default_rng(11)
    cases = []
    for H, W in ((96, 128), (77, 131)):
        n = 30
        masks = np.zeros((n, H, W), dtype=bool)
        for i in range(n):
            y0, x0 = rng.integers(0, H - 12), rng.integers(0, W - 12)
            h, w = rng.integers(6, 40), rng.integers(6, 40)
            blob = rng.random((min(h, H - y0), min(w, W - x0))) < 0.8
            masks[i, y0:y0 + blob.shape[0], x0:x0 + blob.shape[1]] = blob
        masks[3] = False
        masks[5, :, 10:14] = True
        cases.append(dict(name=f"random_{H}x{W}", masks=masks, scores=np.linspace(0.95, 0.55, n)))
        thin = np.zeros((5, H, W), bool)
        thin[:, 20:40, 30:32] = True
        cases.append(dict(name=f"thin_{H}x{W}", masks=thin, scores=np.linspace(0.95, 0.55, n)[:5]))
    # hand-checkable KATs (tests/test_oracle_cleanup.py)
    H, W = 12, 14
    ring = np.zeros((H, W), bool); ring[2:8, 2:8] = True; ring[4:6, 4:6] = False
    sq = np.zeros((H, W), bool); sq[5:10, 6:12] = True
    cases.append(dict(name="kat_ring_square", masks=np.stack([ring, sq]), scores=np.array([0.9, 0.8])))
    slit = np.zeros((H, W), bool); slit[3:9, 3:9] = True; slit[3:6, 5] = False; slit[6, 3] = False
    cases.append(dict(name="kat_slit", masks=slit[None], scores=np.array([0.9])))
    bar = np.zeros((H, W), bool); bar[5:7, 1:13] = True
    blocker = np.zeros((H, W), bool); blocker[3:9, 5:9] = True
    cases.append(dict(name="kat_split", masks=np.stack([blocker, bar]), scores=np.array([0.9, 0.8])))
    cases.append(dict(name="kat_zero_score", masks=np.stack([ring, sq]), scores=np.array([0.9, 0.0])))
    one_col = np.zeros((3, H, W), bool); one_col[:, 2:10, 4] = True
    cases.append(dict(name="kat_one_column", masks=one_col, scores=np.array([0.9, 0.8, 0.7])))
    cases.append(dict(name="kat_all_empty", masks=np.zeros((2, H, W), bool), scores=np.array([0.9, 0.8])))
    edge = np.zeros((H, W), bool); edge[0:4, 0:5] = True
    cases.append(dict(name="kat_edge", masks=edge[None], scores=np.array([0.9])))
    near = np.zeros((H, W), bool); near[5:11, 6:12] = True
    cases.append(dict(name="kat_near_border", masks=near[None], scores=np.array([0.9])))
    cases.append(dict(name="kat_no_masks", masks=np.zeros((0, H, W), bool), scores=np.zeros(0)))
    return cases


def rle_cases() -> List[np.ndarray]:
    rng = np.random.default_rng(0)
    out = []
    for shape in ((7, 5), (16, 16), (3, 40), (33, 65)):
        for p in (0.1, 0.5, 0.9):
            out.append((rng.random(shape) < p).astype(np.uint8))
    out.append(np.zeros((4, 4), np.uint8))
    x = np.zeros((4, 3), np.uint8); x[1:, 0] = 1; x[:2, 1] = 1; x[3, 2] = 1
    out.append(x)
    return out
