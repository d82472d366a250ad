"""Single-forward integration.

The reference runs ``predictor(im)`` twelve times per image (GetInference, GetCounts and
GetMask_Contours for each of the four class keywords, nn_inference.py:343-372, :487-496) and
each call ends in Detectron2's ``fast_rcnn_inference`` -> mask head -> ``mask_rcnn_inference``
(sigmoid + class-channel select into an N x 1 x 28 x 28 tensor) -> ``detector_postprocess``.
Here the network runs ONCE per image and everything after the box head's logits is this
library: score filter / per-class NMS / top-k on the GPU for the whole batch
(``fast_rcnn_inference``, kernel 4), then the mask head's raw N x K x 28 x 28 logits go straight
into the paste kernel, which selects the predicted class's channel and applies the sigmoid
while staging the tile (``uwcv_paste_measure_heads``): neither the probability tensor nor the
N x H x W masks are materialised.

The backbone / RPN / RoI heads themselves are the model's own torch modules (out of scope:
SURVEY.md section 7).  ``SingleForward`` drives a torchvision ``MaskRCNN`` (the stand-in for
Detectron2's R50/R101-FPN, which is not installable here: SURVEY.md 8(c)); any model exposing
the same four stages can be driven the same way.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import torch

from .api import Engine, MeasurementStream, _require_cuda, measure_instances
from .structures import Boxes, Instances


def fast_rcnn_inference(boxes: Sequence[torch.Tensor], scores: Sequence[torch.Tensor],
                        image_shapes: Sequence[Tuple[int, int]], score_thresh: float,
                        nms_thresh: float, topk_per_image: int, device=None):
    """detectron2.modeling.roi_heads.fast_rcnn.fast_rcnn_inference for a list of images, in
    ONE batched kernel sequence.  ``boxes[b]`` is R_b x (K*4) (or R_b x 4), ``scores[b]`` is
    R_b x (K+1) with the background column LAST.  Returns (list of Instances with pred_boxes /
    scores / pred_classes, list of kept proposal-row indices -- positions after Detectron2's
    finite-value filter, i.e. ``filter_inds[:, 0]``), as Detectron2 does."""
    dev = _require_cuda(device if device is not None else
                        (boxes[0].device if len(boxes) and boxes[0].is_cuda else None))
    eng = Engine.get(dev)
    cb, cs, cc, off = [], [], [], [0]
    K = 0
    for bx, sc, (h, w) in zip(boxes, scores, image_shapes):
        bx = bx.to(dev, torch.float32)
        sc = sc.to(dev, torch.float32)
        valid = torch.isfinite(bx).all(dim=1) & torch.isfinite(sc).all(dim=1)
        if not bool(valid.all()):
            bx, sc = bx[valid], sc[valid]
        sc = sc[:, :-1]
        R, K = sc.shape
        nreg = bx.shape[1] // 4
        b4 = bx.reshape(-1, 4)
        b4 = torch.stack((b4[:, 0].clamp(min=0, max=w), b4[:, 1].clamp(min=0, max=h),
                          b4[:, 2].clamp(min=0, max=w), b4[:, 3].clamp(min=0, max=h)), dim=-1)
        b4 = b4.view(R, nreg, 4)
        if nreg == 1:
            b4 = b4.expand(R, K, 4)
        cb.append(b4.reshape(R * K, 4))
        cs.append(sc.reshape(R * K))
        cc.append(torch.arange(K, device=dev, dtype=torch.int64).repeat(R))
        off.append(off[-1] + R * K)
    if not cb:
        return [], []
    cand_boxes = torch.cat(cb).contiguous()
    cand_scores = torch.cat(cs).contiguous()
    cand_cls = torch.cat(cc).contiguous()
    keep, cnt = eng.nms(cand_boxes, cand_scores, cand_cls, off, score_thresh, nms_thresh,
                        topk_per_image, num_classes=K)
    counts = cnt.cpu().tolist()
    results, kept_rows = [], []
    for b, (h, w) in enumerate(image_shapes):
        k = keep[off[b]:off[b] + counts[b]]
        res = Instances((int(h), int(w)))
        res.pred_boxes = Boxes(cand_boxes[k])
        res.scores = cand_scores[k]
        res.pred_classes = cand_cls[k]
        results.append(res)
        kept_rows.append((k - off[b]) // K)        # positions after the finite-value filter (D2)
    return results, kept_rows


class SingleForward:
    """One network forward per image; score filter, NMS, mask selection, sigmoid, paste and
    measurement on the GPU in this library.

        sf = uwcv.SingleForward(model, score_thresh=0.8)      # nn_inference.py:226
        table = sf.measure(images)                            # list of C x H x W float tensors
        rows_of_class_2 = table.for_class(2).reference_rows() # what GetMask_Contours appends

    ``model``: a torchvision ``MaskRCNN`` in eval mode on a CUDA device."""

    def __init__(self, model, score_thresh: float = 0.8, nms_thresh: float = 0.5,
                 detections_per_image: int = 100, device=None):
        self.model = model.eval()
        p = next(model.parameters())
        self.device = _require_cuda(device if device is not None else p.device)
        self.score_thresh = float(score_thresh)
        self.nms_thresh = float(nms_thresh)
        self.topk = int(detections_per_image)

    @torch.no_grad()
    def predict(self, images: Sequence[torch.Tensor]) -> List[Instances]:
        """Raw predictor output per image: ``pred_boxes`` in network-input coordinates
        (``image_size`` = the resized image), ``scores``, ``pred_classes`` (0-based, as
        Detectron2) and ``pred_mask_logits`` (N x (K+1) x 28 x 28 views into one tensor;
        channel 0 is torchvision's background, hence ``mask_channel_offset=1`` downstream)."""
        m = self.model
        images = [im.to(self.device) for im in images]
        ilist, _ = m.transform(images)
        feats = m.backbone(ilist.tensors)
        if isinstance(feats, torch.Tensor):
            feats = OrderedDict([("0", feats)])
        proposals, _ = m.rpn(ilist, feats)
        rh = m.roi_heads
        box_feats = rh.box_head(rh.box_roi_pool(feats, proposals, ilist.image_sizes))
        class_logits, box_regression = rh.box_predictor(box_feats)
        per_image = [int(p.shape[0]) for p in proposals]
        pred_boxes = rh.box_coder.decode(box_regression, proposals)        # R x (K+1) x 4
        pred_scores = torch.softmax(class_logits, -1)                      # background FIRST
        # Detectron2 layout: foreground classes 0..K-1, background column last
        boxes_l = [b[:, 1:].reshape(b.shape[0], (b.shape[1] - 1) * 4)
                   for b in pred_boxes.split(per_image, 0)]
        scores_l = [torch.cat((s[:, 1:], s[:, :1]), dim=1) for s in pred_scores.split(per_image, 0)]
        results, _ = fast_rcnn_inference(boxes_l, scores_l, ilist.image_sizes, self.score_thresh,
                                         self.nms_thresh, self.topk, device=self.device)
        kept = [r.pred_boxes.tensor for r in results]
        counts = [int(k.shape[0]) for k in kept]
        if sum(counts) > 0:
            mask_feats = rh.mask_head(rh.mask_roi_pool(feats, kept, ilist.image_sizes))
            logits = rh.mask_predictor(mask_feats).to(torch.float32).contiguous()
        else:
            logits = torch.zeros((0, 2, 28, 28), dtype=torch.float32, device=self.device)
        lo = 0
        for r, c in zip(results, counts):
            r.set("pred_mask_logits", logits[lo:lo + c])
            lo += c
        return results

    def measure(self, images: Sequence[torch.Tensor], classes_of_interest=None,
                output_size: Optional[Tuple[int, int]] = None, **kw):
        """Measurement table of a batch of same-sized images (``output_size`` defaults to the
        size of the given images: the boxes are rescaled to it as detector_postprocess does)."""
        if output_size is None:
            output_size = tuple(int(v) for v in images[0].shape[-2:])
        inst = self.predict(images)
        return measure_instances(inst, output_size, classes_of_interest,
                                 mask_channel_offset=1, device=self.device, **kw)

    def measure_stream(self, batches, classes_of_interest=None,
                       output_size: Optional[Tuple[int, int]] = None, depth: int = 2, **kw):
        """Generator of tables over an iterable of image batches, two batches in flight: the
        network forward of batch i + 1 is enqueued while batch i is being measured / read."""
        stream = MeasurementStream(self.device, depth=depth)

        def gen():
            for images in batches:
                size = output_size or tuple(int(v) for v in images[0].shape[-2:])
                yield self.predict(images), size

        inflight = []
        for inst, size in gen():
            inflight.append(stream.submit(inst, size, classes_of_interest,
                                          mask_channel_offset=1, **kw))
            if len(inflight) >= depth:
                yield inflight.pop(0).result()
        while inflight:
            yield inflight.pop(0).result()
