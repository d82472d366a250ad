"""Oracle end-to-end drivers (TEST INFRASTRUCTURE): predictor output -> rows, on the CPU,
through the restated Detectron2 glue (oracle.d2) and measurement block (oracle.measure).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import d2
from . import measure as M


def to_oracle_instances(inst) -> d2.Instances:
    """Copy any Instances duck type (CPU tensors) into the oracle's container."""
    o = d2.Instances(tuple(inst.image_size))
    o.pred_boxes = d2.Boxes(inst.pred_boxes.tensor.detach().cpu().clone())
    o.scores = inst.scores.detach().cpu().clone()
    o.pred_classes = inst.pred_classes.detach().cpu().clone()
    o.pred_masks = inst.pred_masks.detach().cpu().clone()
    return o


def postprocess_boxes(inst: d2.Instances, out_size: Tuple[int, int]) -> d2.Instances:
    """detector_postprocess up to (not including) the paste: scale, clip, drop empty."""
    H, W = out_size
    scale_x, scale_y = (W / inst.image_size[1], H / inst.image_size[0])
    res = d2.Instances((H, W), **inst.get_fields())
    boxes = res.pred_boxes.clone()
    res.remove("pred_boxes")
    res.pred_boxes = boxes
    boxes.scale(scale_x, scale_y)
    boxes.clip(res.image_size)
    return res[boxes.nonempty()]


def pack_bits(mask_hw: np.ndarray, row_words: Optional[int] = None) -> np.ndarray:
    """bool H x W -> uint32 H x row_words, bit b of word w = pixel x = 32 w + b."""
    H, W = mask_hw.shape
    nw = (W + 31) // 32
    row_words = row_words or ((nw + 3) // 4 * 4)
    padded = np.zeros((H, row_words * 32), dtype=np.uint8)
    padded[:, :W] = mask_hw
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(H, row_words)


def oracle_windows(res: d2.Instances, thr: float = 0.5):
    """Yield (bool window, y0, x0) per instance through the CPU paste path."""
    H, W = res.image_size
    masks = res.pred_masks[:, 0] if res.pred_masks.dim() == 4 else res.pred_masks
    for i in range(len(res)):
        win, y0, x0 = d2.paste_one_cropped(masks[i].float(), res.pred_boxes.tensor[i], H, W, thr)
        yield win.numpy(), y0, x0


def oracle_table(batch: Sequence, out_size: Tuple[int, int],
                 classes_of_interest: Optional[Sequence[int]] = None, thr: float = 0.5,
                 ppm: float = 0.85, image_idx_offset: int = 0):
    """Per-instance rows for a list of raw predictor outputs (one Instances per image)."""
    I, F = [], []
    for k, inst in enumerate(batch):
        o = to_oracle_instances(inst)
        if classes_of_interest is not None:
            sel = torch.tensor([int(c) in [int(v) for v in classes_of_interest]
                                for c in o.pred_classes], dtype=torch.bool)
            o = o[sel]
        if len(o) == 0:
            continue
        res = postprocess_boxes(o, out_size)
        if len(res) == 0:
            continue
        ri, rf = M.instance_rows(oracle_windows(res, thr), res.pred_classes.numpy(),
                                 res.scores.numpy(), image_idx=image_idx_offset + k,
                                 pixels_per_metric=ppm)
        I.append(ri)
        F.append(rf)
    if not I:
        return (np.zeros((0, len(M.INT_COLUMNS)), np.int64),
                np.zeros((0, len(M.FLOAT_COLUMNS)), np.float64))
    return np.concatenate(I), np.concatenate(F)


def reference_literal_rows(inst, out_size: Tuple[int, int], classes_of_interest: Sequence[int],
                           thr: float = 0.5, literal_paint: bool = False):
    """What the reference script computes for one image and one class keyword:
    predictor post-process (full N x H x W bool) -> GetMask_Contours rows."""
    o = to_oracle_instances(inst)
    res = d2.detector_postprocess(o, out_size[0], out_size[1], thr)
    H, W = out_size
    return M.get_mask_contours((H, W, 3), res.pred_classes.numpy(), res.pred_masks.numpy(),
                               classes_of_interest, literal_paint=literal_paint)
