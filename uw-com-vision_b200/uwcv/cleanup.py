"""Mask clean-up + RLE export.

Replaces the reference's export loop nn_inference.py:319-336 -- ``postprocess_masks``
(:265-306: fill holes, dilate + erode, cut overlaps in score order, empty masks that fall into
several pieces) and ``rle_encoding`` (:253-263: column-major, 1-based (start, length) pairs) --
and the ``R50_flip_.csv`` it writes (``ImageId, EncodedPixels``).

``export_rle`` takes the RAW predictor output (as ``measure_instances`` does), pastes the masks
into bit tiles on the GPU and runs the clean-up and the run extraction there
(``csrc/cleanup.cu``), the ``EncodedPixels`` text included (digits printed by a kernel); the host
only slices the text per instance.  The reference's quirks are kept
(the image is skipped when a score is exactly zero; the instance list is truncated to the number
of image columns holding more than ``min_crys_size`` mask pixels when that number is smaller
than the instance count): see oracle/cleanup.py for the line-by-line reading.
"""
from __future__ import annotations

import csv
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .api import (Engine, MASK_SIDE, _as_box_tensor, _mask_field, _ptr, _require_cuda, _stream_ptr,
                  scale_clip_boxes, tile_words)
from .schema import NUM_FLOAT, NUM_INT


@dataclass
class RleExport:
    """One entry per exported mask, image-major, list order inside an image (the rows of the
    reference's DataFrame).  ``encoded_pixels[k]`` is ``''`` for an emptied mask."""
    image_id: List[str]
    encoded_pixels: List[str]
    image_idx: np.ndarray          # [K] position of the image in the call
    inst_idx: np.ndarray           # [K] index of the instance inside its image (after the
    #                                    empty-box filter of detector_postprocess)
    area: np.ndarray               # [K] pixels of the cleaned mask
    multi_piece: np.ndarray        # [K] bool: emptied because it fell into several pieces

    def __len__(self) -> int:
        return len(self.encoded_pixels)


def _clean_on_workspace(eng: Engine, n: int, H: int, W: int, d_slot, d_inst, counts, skip,
                        min_crys_size: int):
    """Column totals -> per-image limit (the reference's keep_ind quirk, :277-284) -> clean-up
    kernels on the engine's tile workspace.  Returns device (limit, flags, area, run counts)."""
    dev = eng.device
    L = eng.L
    ws = eng._ws
    st = _stream_ptr(dev)
    B = len(counts)
    coltot = torch.zeros((B, W), dtype=torch.int32, device=dev)
    _lib.check(L.uwcv_mask_column_totals(_ptr(ws), ws.numel(), n, W, _ptr(d_slot), _ptr(coltot), st),
               "uwcv_mask_column_totals")
    kcols = (coltot > int(min_crys_size)).sum(dim=1)
    n_img = torch.tensor(counts, dtype=torch.int64, device=dev)
    limit = torch.where(kcols < n_img, kcols, n_img)
    limit = torch.where(torch.tensor(skip, device=dev), torch.zeros_like(limit), limit)
    d_limit = limit.to(torch.int32).contiguous()
    flags = torch.empty(n, dtype=torch.int32, device=dev)
    area = torch.empty(n, dtype=torch.int64, device=dev)
    nruns = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(L.uwcv_clean_masks(_ptr(ws), ws.numel(), n, H, W, _ptr(d_slot), _ptr(d_inst),
                                  _ptr(d_limit), _ptr(flags), _ptr(area), _ptr(nruns), st),
               "uwcv_clean_masks")
    eng.launches += 3
    return d_limit, flags, area, nruns


def postprocess_masks(ori_mask, ori_score, image, min_crys_size: int = 2, *, device=None):
    """The reference's ``postprocess_masks(ori_mask, ori_score, image, min_crys_size=2)``
    (nn_inference.py:265-306) with the same arguments and return value: ``ori_mask`` N x H x W
    bool (Detectron2's pasted ``pred_masks``, numpy or torch), ``ori_score`` N floats, ``image``
    the H x W (x C) image or its shape; returns ``None`` when there is no mask or a score is
    exactly zero (:274), else the list of cleaned H x W uint8 masks (possibly truncated, :277-284).
    The masks are packed into bit tiles on the GPU and cleaned there (``csrc/cleanup.cu``)."""
    dev = _require_cuda(device)
    eng = Engine.get(dev)
    m = torch.as_tensor(np.asarray(ori_mask) if not torch.is_tensor(ori_mask) else ori_mask)
    sc = torch.as_tensor(np.asarray(ori_score) if not torch.is_tensor(ori_score) else ori_score)
    shape = image.shape if hasattr(image, "shape") else tuple(image)
    H, W = int(shape[0]), int(shape[1])
    n = int(m.shape[0]) if m.dim() == 3 else 0
    # ``len(ori_mask) == 0 or ori_score.all() < score_threshold``
    if n == 0 or not bool((sc != 0).all()):
        return None
    if tuple(m.shape[1:]) != (H, W):
        raise ValueError(f"masks {tuple(m.shape)} do not match the image {H} x {W}")
    L = eng.L
    with torch.cuda.device(dev):
        d_m = m.to(dev).to(torch.uint8).contiguous()
        st = _stream_ptr(dev)
        d_boxes = torch.empty((n, 4), dtype=torch.float32, device=dev)
        _lib.check(L.uwcv_mask_pixel_boxes(_ptr(d_m), n, H, W, _ptr(d_boxes), st), "uwcv_mask_pixel_boxes")
        d_slot = torch.zeros(n, dtype=torch.int32, device=dev)
        d_inst = torch.arange(n, dtype=torch.int32, device=dev)
        rows_i = torch.empty((n, NUM_INT), dtype=torch.int64, device=dev)
        rows_f = torch.empty((n, NUM_FLOAT), dtype=torch.float64, device=dev)
        # the layout stage alone (tiles with a margin around every pixel box), then the tile
        # plane is filled from the given masks instead of being pasted
        eng.run(d_m, d_boxes, H, W, image_idx=d_slot, inst_idx=d_inst, rows_i=rows_i, rows_f=rows_f, stages=1,
                n_tile_words=tile_words(d_boxes, H, W))
        eng.check_status()
        ws = eng._ws
        _lib.check(L.uwcv_pack_mask_tiles(_ptr(d_m), n, H, W, _ptr(ws), ws.numel(), st),
                   "uwcv_pack_mask_tiles")
        d_limit, flags, _area, _nruns = _clean_on_workspace(eng, n, H, W, d_slot, d_inst, [n], [False],
                                                            min_crys_size)
        keep = int(d_limit.cpu()[0])
        if keep == 0:
            return []
        # (the workspace is carved for all n instances; the masks beyond the limit are empty)
        out = torch.zeros((n, H, W), dtype=torch.uint8, device=dev)
        _lib.check(L.uwcv_tiles_to_masks(_ptr(ws), ws.numel(), n, H, W, _ptr(out), st),
                   "uwcv_tiles_to_masks")
        eng.launches += 3
        host = out[:keep].cpu().numpy()
    return [host[i] for i in range(keep)]


def export_rle(instances, output_size: Optional[Tuple[int, int]] = None,
               names: Optional[Sequence[str]] = None, *, mask_threshold: float = 0.5,
               min_crys_size: int = 2, mask_channel_offset: int = 0, device=None) -> RleExport:
    """Clean the masks of a batch of images and run-length encode them.

    ``instances``: one ``Instances`` per image (or a single one) with the raw predictor output,
    in Detectron2's order (score descending) -- the order decides who keeps an overlap.
    ``names``: image file names (``.tif`` is stripped for ``ImageId``, :331); default "0", "1", ...
    """
    single = not isinstance(instances, (list, tuple))
    batch = [instances] if single else list(instances)
    dev = _require_cuda(device)
    eng = Engine.get(dev)
    names = [str(k) for k in range(len(batch))] if names is None else list(names)
    if len(names) != len(batch):
        raise ValueError("one name per image")
    empty = RleExport([], [], np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64),
                      np.zeros(0, bool))
    if not batch:
        return empty
    H, W = (int(output_size[0]), int(output_size[1])) if output_size is not None else \
        tuple(int(v) for v in batch[0].image_size)
    bl, ml, cl, slot_l, inst_l, counts, skip = [], [], [], [], [], [], []
    logits, channels = False, 1
    for k, inst in enumerate(batch):
        masks, logits = _mask_field(inst)
        channels = int(masks.shape[1])
        b, keep = scale_clip_boxes(_as_box_tensor(inst.pred_boxes), inst.image_size, (H, W))
        b, masks = b[keep], masks[keep]
        scores = inst.scores[keep]
        nk = int(b.shape[0])
        bl.append(b.cpu())
        ml.append(masks.to(torch.float32).reshape(nk, channels * MASK_SIDE * MASK_SIDE).cpu())
        cl.append(inst.pred_classes[keep].to(torch.int64).cpu())
        slot_l.append(torch.full((nk,), k, dtype=torch.int32))
        inst_l.append(torch.arange(nk, dtype=torch.int32))
        counts.append(nk)
        # ``ori_score.all() < 0.5`` (:274): a bool compared with 0.5 -- true iff some score is 0
        skip.append(nk == 0 or bool((scores == 0).any()))
    n = sum(counts)
    if n == 0:
        return empty
    L = eng.L
    B = len(batch)
    with torch.cuda.device(dev):
        boxes = torch.cat(bl)
        d_boxes = boxes.contiguous().to(dev)
        d_masks = torch.cat(ml).contiguous().to(dev)
        d_cls = torch.cat(cl).to(dev)
        d_slot = torch.cat(slot_l).to(dev)
        d_inst = torch.cat(inst_l).to(dev)
        rows_i = torch.empty((n, NUM_INT), dtype=torch.int64, device=dev)
        rows_f = torch.empty((n, NUM_FLOAT), dtype=torch.float64, device=dev)
        # layout + paste into bit tiles (no full-frame planes, no border trace)
        eng.run(d_masks, d_boxes, H, W, image_idx=d_slot, inst_idx=d_inst, classes=d_cls,
                threshold=mask_threshold, rows_i=rows_i, rows_f=rows_f, stages=3,
                n_tile_words=tile_words(boxes, H, W), mask_channels=channels,
                channel_offset=mask_channel_offset if channels > 1 else 0, logits=logits)
        eng.check_status()
        ws = eng._ws
        st = _stream_ptr(dev)
        d_limit, flags, area, nruns = _clean_on_workspace(eng, n, H, W, d_slot, d_inst, counts, skip,
                                                          min_crys_size)
        run_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(nruns, 0, out=run_off[1:])
        total = int(run_off[-1].item())
        runs = torch.empty((max(total, 1), 2), dtype=torch.int64, device=dev)
        _lib.check(L.uwcv_rle_write(_ptr(ws), ws.numel(), n, H, W, _ptr(run_off), _ptr(runs), st),
                   "uwcv_rle_write")
        eng.launches += 1
        # the text on the device: characters per (merged) run -> offsets -> digits
        run_inst = torch.repeat_interleave(torch.arange(n, dtype=torch.int32, device=dev), nruns,
                                           output_size=total)
        chars = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
        _lib.check(L.uwcv_rle_text_prep(total, _ptr(runs), _ptr(run_inst), _ptr(chars), st),
                   "uwcv_rle_text_prep")
        toff = torch.zeros(total + 1, dtype=torch.int64, device=dev)
        if total:
            torch.cumsum(chars[:total], 0, out=toff[1:])
        inst_text = toff[run_off]                              # [n + 1] first character of every instance
        nchar = int(toff[-1].item())
        text = torch.empty(max(nchar, 1), dtype=torch.uint8, device=dev)
        _lib.check(L.uwcv_rle_text_write(total, _ptr(runs), _ptr(run_inst), _ptr(toff), _ptr(text), st),
                   "uwcv_rle_text_write")
        eng.launches += 3
        h_text = text[:nchar].cpu().numpy().tobytes()
        h_bounds = inst_text.cpu().tolist()
        h_flags = flags.cpu().numpy()
        h_area = area.cpu().numpy()
        h_limit = d_limit.cpu().numpy()
    slot = torch.cat(slot_l).numpy()
    idx = torch.cat(inst_l).numpy()
    out = RleExport([], [], None, None, None, None)
    kr = np.flatnonzero(idx < h_limit[slot])              # instances that are part of the output
    ids = [nm.replace('.tif', '') for nm in names]
    for i, b in zip(kr.tolist(), slot[kr].tolist()):
        out.image_id.append(ids[b])
        lo_c, hi_c = h_bounds[i], h_bounds[i + 1]
        out.encoded_pixels.append(h_text[lo_c: hi_c - 1].decode("ascii") if hi_c > lo_c else "")
    out.image_idx = slot[kr].astype(np.int64)
    out.inst_idx = idx[kr].astype(np.int64)
    out.area = h_area[kr]
    out.multi_piece = (h_flags[kr] & 1).astype(bool)
    return out


def write_rle_csv(path: str, export: RleExport) -> None:
    """``pd.DataFrame({"ImageId": ..., "EncodedPixels": ...}).to_csv(path, index=False)`` (:335-336)."""
    with open(path, "w", newline="") as f:
        w = csv.writer(f, lineterminator="\n")            # pandas' line terminator
        w.writerow(["ImageId", "EncodedPixels"])
        for a, b in zip(export.image_id, export.encoded_pixels):
            w.writerow([a, b])
