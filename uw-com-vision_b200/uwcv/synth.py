"""Seeded synthetic predictor outputs for the BASELINE.json configs.

No datasets or weights exist offline, so the measured workloads are synthetic:
``blob_instances`` draws Mask R-CNN-shaped predictions (boxes, unique scores, classes,
28 x 28 mask probabilities shaped like soft blobs with holes / speckle);
``clustered_candidates`` draws the dense candidate list of the high-density config
(stresses NMS).  Everything is generated on the CPU from ``torch.Generator`` seeds so
that the oracle and the CUDA path see identical bytes.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn.functional as F

from .structures import Boxes, Instances

MASK_SIDE = 28


def _unique_scores(n: int, lo: float, hi: float, g: torch.Generator) -> torch.Tensor:
    """float32 scores with no ties (rejection, not an additive tie-breaker: SURVEY.md H6)."""
    s = torch.rand(n, generator=g) * (hi - lo) + lo
    for _ in range(64):
        u, inv, cnt = torch.unique(s, return_inverse=True, return_counts=True)
        if u.numel() == n:
            break
        dup = cnt[inv] > 1
        s[dup] = torch.rand(int(dup.sum()), generator=g) * (hi - lo) + lo
    return s.to(torch.float32)


def blob_probs(n: int, g: torch.Generator, speckle_frac: float = 0.1) -> torch.Tensor:
    """n x 28 x 28 probabilities: sigmoid(8 (0.8 - r_ellipse) + low-pass noise), a random
    ellipse orientation / aspect per instance; a fraction gets saturated 0/1 speckle and
    exact 0.5 plateaus (threshold ties)."""
    lin = (torch.arange(MASK_SIDE, dtype=torch.float32) + 0.5) / MASK_SIDE * 2 - 1
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    th = torch.rand(n, 1, 1, generator=g) * math.pi
    ax = 0.55 + 0.45 * torch.rand(n, 1, 1, generator=g)
    ay = ax * (0.35 + 0.65 * torch.rand(n, 1, 1, generator=g))
    cx = (torch.rand(n, 1, 1, generator=g) - 0.5) * 0.3
    cy = (torch.rand(n, 1, 1, generator=g) - 0.5) * 0.3
    u = (xx - cx) * torch.cos(th) + (yy - cy) * torch.sin(th)
    v = -(xx - cx) * torch.sin(th) + (yy - cy) * torch.cos(th)
    r = torch.sqrt((u / ax) ** 2 + (v / ay) ** 2)
    noise = torch.randn(n, 1, MASK_SIDE, MASK_SIDE, generator=g)
    k = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0])
    k2 = (k[:, None] * k[None, :] / 256.0)[None, None]
    noise = F.conv2d(noise, k2, padding=2)[:, 0] * 2.5
    p = torch.sigmoid(8.0 * (0.8 - r) + noise)
    ns = int(n * speckle_frac)
    if ns:
        idx = torch.randperm(n, generator=g)[:ns]
        q = p[idx]
        sat = torch.rand(q.shape, generator=g)
        q = torch.where(sat < 0.15, torch.zeros_like(q), q)
        q = torch.where(sat > 0.85, torch.ones_like(q), q)
        q = torch.where((sat > 0.48) & (sat < 0.52), torch.full_like(q, 0.5), q)
        p[idx] = q
    return p.to(torch.float32).contiguous()


def blob_instances(image_index: int, n: int, H: int, W: int, seed: int = 1234,
                   size_range: Tuple[float, float] = (16.0, 128.0),
                   num_classes: int = 4) -> Instances:
    """One image of config-2 style predictions, network input size == output size."""
    g = torch.Generator().manual_seed(seed + image_index)
    cxy = torch.rand(n, 2, generator=g) * torch.tensor([W, H], dtype=torch.float32)
    lo, hi = math.log(size_range[0]), math.log(size_range[1])
    wh = torch.exp(torch.rand(n, 2, generator=g) * (hi - lo) + lo)
    b = torch.cat([cxy - wh / 2, cxy + wh / 2], dim=1)
    b[:, 0::2] = b[:, 0::2].clamp(0, W)
    b[:, 1::2] = b[:, 1::2].clamp(0, H)
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    b = b[keep].to(torch.float32)
    m = int(b.shape[0])
    inst = Instances((H, W))
    inst.pred_boxes = Boxes(b)
    inst.scores = _unique_scores(m, 0.05, 1.0, g)
    inst.pred_classes = torch.randint(0, num_classes, (m,), generator=g, dtype=torch.int64)
    inst.pred_masks = blob_probs(m, g)[:, None]
    return inst


def blob_batch(n_images: int, n_per_image: int, H: int, W: int, seed: int = 1234,
               first_image: int = 0, stride: int = 1, **kw) -> List[Instances]:
    return [blob_instances(first_image + i * stride, n_per_image, H, W, seed, **kw)
            for i in range(n_images)]


def clustered_candidates(n_seeds: int, H: int, W: int, seed: int = 99, n_clusters: int = 50,
                         siblings: int = 4, num_classes: int = 4):
    """High-density config: ``n_seeds`` boxes (sides U[8, 48]) around Gaussian clusters, each
    expanded into jittered sibling candidates with IoU > 0.5 among siblings.
    Returns (boxes [R, 4], scores [R], classes [R])."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.rand(n_clusters, 2, generator=g) * torch.tensor([W, H], dtype=torch.float32)
    which = torch.randint(0, n_clusters, (n_seeds,), generator=g)
    c = centres[which] + torch.randn(n_seeds, 2, generator=g) * 40.0
    wh = torch.rand(n_seeds, 2, generator=g) * 40.0 + 8.0
    cls = torch.randint(0, num_classes, (n_seeds,), generator=g, dtype=torch.int64)
    c = c[:, None, :] + (torch.rand(n_seeds, siblings, 2, generator=g) - 0.5) * 0.2 * wh[:, None, :]
    whs = wh[:, None, :] * (1.0 + (torch.rand(n_seeds, siblings, 2, generator=g) - 0.5) * 0.2)
    b = torch.cat([c - whs / 2, c + whs / 2], dim=2).reshape(-1, 4)
    b[:, 0::2] = b[:, 0::2].clamp(0, W)
    b[:, 1::2] = b[:, 1::2].clamp(0, H)
    cls = cls[:, None].expand(n_seeds, siblings).reshape(-1).contiguous()
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    b, cls = b[keep].to(torch.float32).contiguous(), cls[keep].contiguous()
    s = _unique_scores(int(b.shape[0]), 0.0, 1.0, g)
    return b, s, cls
