"""Oracle restatement of the Detectron2 glue on the hot path (TEST INFRASTRUCTURE).

Detectron2 is not vendored in /root/reference and cannot be installed offline
(SURVEY.md section 8(c)); the reference installs HEAD of
github.com/facebookresearch/detectron2 (COLAB_PORT.py:4).  These functions
restate the published upstream algorithms the reference reaches through
``predictor(im)`` (nn_inference.py:222-227, :372):

* ``Boxes`` / ``Instances``       -- detectron2/structures/{boxes,instances}.py
* ``detector_postprocess``        -- detectron2/modeling/postprocessing.py
* ``paste_masks_in_image``        -- detectron2/layers/mask_ops.py
* ``fast_rcnn_inference_single_image`` / ``fast_rcnn_inference`` -- detectron2/modeling/roi_heads/fast_rcnn.py
* ``mask_rcnn_inference``         -- detectron2/modeling/roi_heads/mask_head.py
* ``batched_nms_vanilla``         -- torchvision/ops/boxes.py:_batched_nms_vanilla

The arithmetic itself is delegated to the libraries the reference runs on
(``torch.nn.functional.grid_sample``, ``torchvision.ops.nms``).
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, List, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

BYTES_PER_FLOAT = 4
GPU_MEM_LIMIT = 1024 ** 3  # upstream mask_ops.py: 1 GB memory limit


def assert_cpu_capability() -> str:
    """The bit-exact paste recipe is pinned to the vectorised ATen CPU kernel.

    SURVEY.md section 8(c): AVX2 and AVX512 are bit-identical, the scalar
    DEFAULT path (no FMA contraction) is not.
    """
    cap = torch.backends.cpu.get_cpu_capability()
    if cap not in ("AVX2", "AVX512"):
        raise RuntimeError(
            f"oracle paste is pinned to the AVX2/AVX512 grid_sample kernel, got {cap!r}")
    return cap


class Boxes:
    """N x 4 float32 XYXY boxes (detectron2/structures/boxes.py)."""

    def __init__(self, tensor: torch.Tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        else:
            tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self) -> "Boxes":
        return Boxes(self.tensor.clone())

    def to(self, *args, **kwargs) -> "Boxes":
        return Boxes(self.tensor.to(*args, **kwargs))

    def area(self) -> torch.Tensor:
        box = self.tensor
        return (box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1])

    def clip(self, box_size: Tuple[int, int]) -> None:
        assert torch.isfinite(self.tensor).all(), "Box tensor contains infinite or NaN!"
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold: float = 0.0) -> torch.Tensor:
        box = self.tensor
        widths = box[:, 2] - box[:, 0]
        heights = box[:, 3] - box[:, 1]
        return (widths > threshold) & (heights > threshold)

    def scale(self, scale_x: float, scale_y: float) -> None:
        self.tensor[:, 0::2] *= scale_x
        self.tensor[:, 1::2] *= scale_y

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2
        return Boxes(b)

    def __len__(self) -> int:
        return self.tensor.shape[0]

    @property
    def device(self):
        return self.tensor.device


class Instances:
    """Field container (detectron2/structures/instances.py).

    The reference reads ``.pred_classes``, ``.pred_masks``, ``.scores`` and also
    ``._fields[...]`` directly (nn_inference.py:326-327, :357, :375-376).
    """

    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        if name == "_fields" or name not in self._fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return self._fields[name]

    def set(self, name: str, value: Any) -> None:
        data_len = len(value)
        if len(self._fields):
            assert len(self) == data_len, (
                f"Adding a field of length {data_len} to a Instances of length {len(self)}")
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def remove(self, name: str) -> None:
        del self._fields[name]

    def get(self, name: str) -> Any:
        return self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, *args: Any, **kwargs: Any) -> "Instances":
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        return ret

    def __getitem__(self, item: Union[int, slice, torch.Tensor]) -> "Instances":
        if type(item) is int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances index out of range!")
            item = slice(item, None, len(self))
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self) -> int:
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")


# --------------------------------------------------------------------------
# mask paste  (UPSTREAM detectron2/layers/mask_ops.py)
# --------------------------------------------------------------------------

def _do_paste_mask(masks, boxes, img_h: int, img_w: int, skip_empty: bool = True):
    """masks N x 1 x Hm x Wm, boxes N x 4 -> (N x h x w float, spatial slices)."""
    device = masks.device
    if skip_empty:
        x0_int, y0_int = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0).to(
            dtype=torch.int32)
        x1_int = torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w).to(dtype=torch.int32)
        y1_int = torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h).to(dtype=torch.int32)
    else:
        x0_int, y0_int = 0, 0
        x1_int, y1_int = img_w, img_h
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)  # each is Nx1

    N = masks.shape[0]

    img_y = torch.arange(y0_int, y1_int, device=device, dtype=torch.float32) + 0.5
    img_x = torch.arange(x0_int, x1_int, device=device, dtype=torch.float32) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    # img_x, img_y have shapes (N, w), (N, h)

    gx = img_x[:, None, :].expand(N, img_y.size(1), img_x.size(1))
    gy = img_y[:, :, None].expand(N, img_y.size(1), img_x.size(1))
    grid = torch.stack([gx, gy], dim=3)

    if not masks.dtype.is_floating_point:
        masks = masks.float()
    img_masks = F.grid_sample(masks, grid.to(masks.dtype), align_corners=False)

    if skip_empty:
        return img_masks[:, 0], (slice(int(y0_int), int(y1_int)), slice(int(x0_int), int(x1_int)))
    return img_masks[:, 0], ()


def paste_masks_in_image(masks: torch.Tensor, boxes, image_shape: Tuple[int, int],
                         threshold: float = 0.5) -> torch.Tensor:
    """N x Hm x Wm probabilities + N x 4 boxes -> N x H x W bool (CPU: one instance at a time)."""
    assert masks.shape[-1] == masks.shape[-2], "Only square mask predictions are supported"
    N = len(masks)
    if N == 0:
        return masks.new_empty((0,) + tuple(image_shape), dtype=torch.uint8)
    if not isinstance(boxes, torch.Tensor):
        boxes = boxes.tensor
    device = boxes.device
    assert len(boxes) == N, boxes.shape

    img_h, img_w = image_shape

    if device.type == "cpu":
        num_chunks = N
    else:
        num_chunks = int(np.ceil(N * int(img_h) * int(img_w) * BYTES_PER_FLOAT / GPU_MEM_LIMIT))
        assert num_chunks <= N, "Default GPU_MEM_LIMIT in mask_ops.py is too small; try increasing it"
    chunks = torch.chunk(torch.arange(N, device=device), num_chunks)

    img_masks = torch.zeros(
        N, img_h, img_w, device=device, dtype=torch.bool if threshold >= 0 else torch.uint8)
    for inds in chunks:
        masks_chunk, spatial_inds = _do_paste_mask(
            masks[inds, None, :, :], boxes[inds], img_h, img_w, skip_empty=device.type == "cpu")
        if threshold >= 0:
            masks_chunk = (masks_chunk >= threshold).to(dtype=torch.bool)
        else:
            masks_chunk = (masks_chunk * 255).to(dtype=torch.uint8)
        img_masks[(inds,) + spatial_inds] = masks_chunk
    return img_masks


def paste_one_cropped(mask: torch.Tensor, box: torch.Tensor, img_h: int, img_w: int,
                      threshold: float = 0.5):
    """One instance through the CPU (skip_empty) path, returned as its window only.

    Same arithmetic as ``paste_masks_in_image`` for that instance (the result of
    ``grid_sample`` is independent of batching: SURVEY.md section 8(c) [PROBE]);
    avoids materialising N x H x W for the full-size configs.
    Returns (bool h x w tensor, y0, x0).
    """
    chunk, (ys, xs) = _do_paste_mask(mask[None, None], box[None], img_h, img_w, skip_empty=True)
    return (chunk[0] >= threshold), ys.start, xs.start


def detector_postprocess(results: Instances, output_height: int, output_width: int,
                         mask_threshold: float = 0.5) -> Instances:
    """UPSTREAM detectron2/modeling/postprocessing.py::detector_postprocess."""
    new_size = (output_height, output_width)
    scale_x, scale_y = (output_width / results.image_size[1],
                        output_height / results.image_size[0])
    results = Instances(new_size, **results.get_fields())

    if results.has("pred_boxes"):
        output_boxes = results.pred_boxes
    elif results.has("proposal_boxes"):
        output_boxes = results.proposal_boxes
    else:
        output_boxes = None
    assert output_boxes is not None, "Predictions must contain boxes!"

    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)

    results = results[output_boxes.nonempty()]

    if results.has("pred_masks"):
        # ROIMasks(pred_masks[:, 0]).to_bitmasks(boxes, H, W, thr).tensor
        results.pred_masks = paste_masks_in_image(
            results.pred_masks[:, 0, :, :], results.pred_boxes.tensor,
            (output_height, output_width), threshold=mask_threshold)
    return results


# --------------------------------------------------------------------------
# box filter + per-class NMS  (UPSTREAM fast_rcnn.py / torchvision boxes.py)
# --------------------------------------------------------------------------

def batched_nms_vanilla(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor,
                        iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.boxes._batched_nms_vanilla: per-class NMS on the original
    coordinates, keep list sorted by score descending."""
    import torchvision
    keep_mask = torch.zeros_like(scores, dtype=torch.bool)
    for class_id in torch.unique(idxs):
        curr_indices = torch.where(idxs == class_id)[0]
        curr_keep_indices = torchvision.ops.nms(boxes[curr_indices], scores[curr_indices],
                                                iou_threshold)
        keep_mask[curr_indices[curr_keep_indices]] = True
    keep_indices = torch.where(keep_mask)[0]
    return keep_indices[scores[keep_indices].sort(descending=True, stable=True)[1]]


def fast_rcnn_inference_single_image(boxes: torch.Tensor, scores: torch.Tensor,
                                     image_shape: Tuple[int, int], score_thresh: float,
                                     nms_thresh: float, topk_per_image: int):
    """UPSTREAM fast_rcnn.py::fast_rcnn_inference_single_image.

    boxes  R x (K*4) (or R x 4 class-agnostic), scores R x (K+1) (last col = background).
    Returns (Instances with pred_boxes/scores/pred_classes, kept row indices R').
    """
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]

    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    boxes = Boxes(boxes.reshape(-1, 4))
    boxes.clip(image_shape)
    boxes = boxes.tensor.view(-1, num_bbox_reg_classes, 4)  # R x C x 4

    filter_mask = scores > score_thresh  # R x K
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]

    keep = batched_nms_vanilla(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]

    result = Instances(image_shape)
    result.pred_boxes = Boxes(boxes)
    result.scores = scores
    result.pred_classes = filter_inds[:, 1]
    return result, filter_inds[:, 0]


def fast_rcnn_inference(boxes, scores, image_shapes, score_thresh: float, nms_thresh: float,
                        topk_per_image: int):
    """UPSTREAM fast_rcnn.py::fast_rcnn_inference: the per-image function over a batch."""
    out = [fast_rcnn_inference_single_image(b, s, shape, score_thresh, nms_thresh, topk_per_image)
           for b, s, shape in zip(boxes, scores, image_shapes)]
    return [x[0] for x in out], [x[1] for x in out]


def mask_rcnn_inference(pred_mask_logits: torch.Tensor, pred_instances: List["Instances"]) -> None:
    """UPSTREAM mask_head.py::mask_rcnn_inference: select the predicted class's channel of the
    N x K x Hm x Wm mask logits (channel 0 for a class-agnostic head), sigmoid, and attach the
    N x 1 x Hm x Wm probabilities to each image's Instances as ``pred_masks``."""
    cls_agnostic_mask = pred_mask_logits.size(1) == 1
    if cls_agnostic_mask:
        mask_probs_pred = pred_mask_logits.sigmoid()
    else:
        num_masks = pred_mask_logits.shape[0]
        class_pred = torch.cat([i.pred_classes for i in pred_instances])
        indices = torch.arange(num_masks, device=class_pred.device)
        mask_probs_pred = pred_mask_logits[indices, class_pred][:, None].sigmoid()
    num_boxes_per_image = [len(i) for i in pred_instances]
    mask_probs_pred = mask_probs_pred.split(num_boxes_per_image, dim=0)
    for prob, instances in zip(mask_probs_pred, pred_instances):
        instances.pred_masks = prob


# --------------------------------------------------------------------------
# Scalar restatement of the paste arithmetic (numpy float32, small cases only).
# SURVEY.md section 8(c) "bit-exact scalar recipe".  Self-tested against
# F.grid_sample in tests/test_oracle_paste.py so a torch upgrade that changes
# the CPU kernel is detected.
# --------------------------------------------------------------------------

def paste_scalar_recipe(mask: np.ndarray, box: np.ndarray, ys: np.ndarray, xs: np.ndarray,
                        threshold: float = 0.5) -> Tuple[np.ndarray, np.ndarray]:
    """Return (float32 sampled values, bool mask) on pixel rows ``ys`` x cols ``xs``."""
    f32 = np.float32
    m = np.ascontiguousarray(mask, dtype=f32)
    M = m.shape[0]
    x0, y0, x1, y1 = [f32(v) for v in box]
    half = f32(M / 2.0)

    def coords(p, lo, hi):
        with np.errstate(all="ignore"):
            g = ((p.astype(f32) + f32(0.5)) - lo) / (hi - lo) * f32(2) - f32(1)
            t = g + f32(1)
            # fmaf(t, M/2, -0.5): exact product in float64 (24b x 24b fits), one rounding
            i = (t.astype(np.float64) * np.float64(half) - 0.5).astype(f32)
            fl = np.floor(i)
            w = i - fl
            e = f32(1) - w
        return fl, w, e

    fx, w, e = coords(xs, x0, x1)
    fy, n, s = coords(ys, y0, y1)

    def tap(fyv, fxv):
        ok = (fyv[:, None] >= 0) & (fyv[:, None] <= M - 1) & (fxv[None, :] >= 0) & (fxv[None, :] <= M - 1)
        yi = np.where(np.isfinite(fyv), np.clip(fyv, 0, M - 1), 0).astype(np.int64)
        xi = np.where(np.isfinite(fxv), np.clip(fxv, 0, M - 1), 0).astype(np.int64)
        return np.where(ok, m[yi[:, None], xi[None, :]], f32(0))

    with np.errstate(all="ignore"):
        nw = s[:, None] * e[None, :]
        ne = s[:, None] * w[None, :]
        sw = n[:, None] * e[None, :]
        se = n[:, None] * w[None, :]
        v_nw = tap(fy, fx)
        v_ne = tap(fy, fx + 1)
        v_sw = tap(fy + 1, fx)
        v_se = tap(fy + 1, fx + 1)

        def fma(a, b, c):  # single rounding: float32 products are exact in float64
            return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

        out = v_nw * nw
        out = fma(v_ne, ne, out)
        out = fma(v_sw, sw, out)
        out = fma(v_se, se, out)
    return out, out >= f32(threshold)
