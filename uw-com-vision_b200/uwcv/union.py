"""Union / connected-component mode: the reference-literal ``GetMask_Contours``.

nn_inference.py:394-459 ORs the masks of the class of interest into ONE image and measures
every external contour of that union (contourArea >= 100), left to right: touching instances
merge and an instance with several blobs yields several rows.  ``measure_union`` reproduces
those rows; the pixel work (paste, OR, border following, descriptors) runs in libuwcv.so, the
host only groups instances by box overlap (N x 4 integers) and orders the small row table.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .api import (Engine, MASK_SIDE, _as_box_tensor, _gather_fields, _ptr, _require_cuda,
                  _stream_ptr, scale_clip_boxes, tile_words)
from .schema import CSV_COLUMNS, ICOL, NUM_FLOAT, NUM_INT

UNION_INT_COLUMNS = ("image_idx", "group", "start_x", "start_y", "brect_x", "brect_y", "brect_w",
                     "brect_h", "n_points", "valid")
UNION_FLOAT_COLUMNS = ("contour_area", "perimeter", "rect_cx", "rect_cy", "rect_w", "rect_h",
                       "rect_angle", "Feret", "Aspect_Ratio", "Roundness", "Circularity",
                       "Sphericity", "Length", "Width", "CircularED", "Chords")


@dataclass
class UnionTable:
    """One row per external contour of the per-image union masks, in the reference's order
    (image-major; inside an image a stable left-to-right sort on cv2.boundingRect x)."""
    ints: np.ndarray       # [K, 10] int64, UNION_INT_COLUMNS
    floats: np.ndarray     # [K, 16] float64, UNION_FLOAT_COLUMNS

    def __len__(self) -> int:
        return self.ints.shape[0]

    def reference_rows(self, image_idx: Optional[int] = None, min_contour_area: float = 100.0
                       ) -> np.ndarray:
        """K x 9 rows in CSV column order (nn_inference.py:561) after the area cut (:412)."""
        ok = self.floats[:, 0] >= min_contour_area
        if image_idx is not None:
            ok &= self.ints[:, 0] == image_idx
        return self.floats[ok][:, 7:16]


def _group_by_overlap(img: np.ndarray, bbox: np.ndarray, valid: np.ndarray):
    """Connected groups of instances (per image) whose 1-pixel-dilated pixel boxes overlap.
    Returns (member_group [N] int32, list of (image, x0, y0, x1, y1) per group)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    n = len(img)
    member = np.full(n, -1, dtype=np.int32)
    groups = []
    for b in np.unique(img[valid]):
        idx = np.flatnonzero(valid & (img == b))
        x0, y0, x1, y1 = (bbox[idx, k].astype(np.int64) for k in range(4))
        # boxes [x0-1, x1+1] x [y0-1, y1+1] intersect  <=>  8-adjacent or overlapping pixel boxes
        ov = (x0[:, None] - 1 <= x1[None, :] + 1) & (x0[None, :] - 1 <= x1[:, None] + 1) & \
             (y0[:, None] - 1 <= y1[None, :] + 1) & (y0[None, :] - 1 <= y1[:, None] + 1)
        r, c = np.nonzero(ov)
        ncomp, lab = connected_components(coo_matrix((np.ones(len(r), np.int8), (r, c)),
                                                     shape=(len(idx), len(idx))), directed=False)
        base = len(groups)
        member[idx] = base + lab
        for g in range(ncomp):
            m = lab == g
            groups.append((int(b), int(x0[m].min()), int(y0[m].min()), int(x1[m].max()),
                           int(y1[m].max())))
    return member, groups


def measure_union(instances, output_size: Optional[Tuple[int, int]] = None,
                  classes_of_interest: Optional[Sequence[int]] = None, *,
                  mask_threshold: float = 0.5, pixels_per_metric: float = 0.85,
                  image_idx_offset: int = 0, device=None) -> UnionTable:
    """Reference-literal measurement rows of one image or a batch (one union per image over the
    instances whose class is in ``classes_of_interest``; all classes when None).
    Instances are raw predictor outputs exactly as for ``measure_instances``."""
    single = not isinstance(instances, (list, tuple))
    batch = [instances] if single else list(instances)
    dev = _require_cuda(device)
    eng = Engine.get(dev)
    empty = UnionTable(np.zeros((0, 10), np.int64), np.zeros((0, 16), np.float64))
    if not batch:
        return empty
    H, W = (int(output_size[0]), int(output_size[1])) if output_size is not None else \
        tuple(int(v) for v in batch[0].image_size)
    bl, ml, il = [], [], []
    for k, inst in enumerate(batch):
        boxes, _scores, _classes, masks = _gather_fields(inst, classes_of_interest)
        b, keep = scale_clip_boxes(boxes, inst.image_size, (H, W))
        b, masks = b[keep], masks[keep]
        bl.append(b.cpu())
        ml.append(masks.to(torch.float32).reshape(-1, MASK_SIDE, MASK_SIDE).cpu())
        il.append(torch.full((int(b.shape[0]),), image_idx_offset + k, dtype=torch.int32))
    boxes = torch.cat(bl)
    n = int(boxes.shape[0])
    if n == 0:
        return empty
    L = eng.L
    with torch.cuda.device(dev):
        d_boxes = boxes.contiguous().to(dev)
        d_masks = torch.cat(ml).contiguous().to(dev)
        d_img = torch.cat(il).to(dev)
        rows_i = torch.empty((n, NUM_INT), dtype=torch.int64, device=dev)
        rows_f = torch.empty((n, NUM_FLOAT), dtype=torch.float64, device=dev)
        # layout + paste (cropped: no full-frame planes), no per-instance contour pass
        eng.run(d_masks, d_boxes, H, W, image_idx=d_img, threshold=mask_threshold,
                pixels_per_metric=pixels_per_metric, rows_i=rows_i, rows_f=rows_f, stages=3,
                n_tile_words=tile_words(boxes, H, W))
        hi = rows_i.cpu().numpy()
        eng.check_status()
        valid = hi[:, ICOL["valid"]] == 1
        if not valid.any():
            return empty
        bbox = hi[:, [ICOL["bbox_x0"], ICOL["bbox_y0"], ICOL["bbox_x1"], ICOL["bbox_y1"]]]
        member, groups = _group_by_overlap(hi[:, ICOL["image_idx"]], bbox, valid)
        G = len(groups)
        gdesc = np.zeros(G, dtype=np.dtype([("wx0", "<i4"), ("y0", "<i4"), ("tw", "<i4"),
                                            ("th", "<i4"), ("word_off", "<i8"), ("res", "<i8")]))
        gimg = np.zeros(G, dtype=np.int32)
        off = 0
        for g, (b, x0, y0, x1, y1) in enumerate(groups):
            wx0 = x0 >> 5
            tw = (x1 >> 5) - wx0 + 1
            th = y1 - y0 + 1
            gdesc[g] = (wx0, y0, tw, th, off, 0)
            gimg[g] = b
            off += tw * th
        gwords = (off + 3) & ~3
        d_member = torch.from_numpy(member).to(dev)
        d_gdesc = torch.from_numpy(gdesc.view(np.uint8).copy()).to(dev)
        d_gimg = torch.from_numpy(gimg).to(dev)
        gplanes = torch.empty(3 * gwords, dtype=torch.int32, device=dev)
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
        rec_cap = 2 * int(valid.sum()) + 256
        ext_cap = 4 * int(gdesc["th"].sum()) + 1024
        ws = eng._ws
        for _attempt in range(3):
            rec_ws = torch.empty(L.uwcv_union_workspace_bytes(rec_cap, ext_cap), dtype=torch.uint8,
                                 device=dev)
            u_i = torch.zeros((rec_cap, 10), dtype=torch.int64, device=dev)
            u_f = torch.zeros((rec_cap, 16), dtype=torch.float64, device=dev)
            rc = L.uwcv_union_measure(_ptr(ws), ws.numel(), n, _ptr(d_member), _ptr(d_gdesc),
                                      _ptr(d_gimg), G, _ptr(gplanes), gwords, _ptr(rec_ws),
                                      rec_ws.numel(), rec_cap, ext_cap, float(pixels_per_metric),
                                      _ptr(u_i), _ptr(u_f), _ptr(counters), _stream_ptr(dev))
            _lib.check(rc, "uwcv_union_measure")
            eng.launches += 4
            c = counters.cpu().tolist()
            if c[1] == 0:
                break
            rec_cap = max(rec_cap, int(c[0]) + 64)          # exact needs reported by the device
            ext_cap = max(ext_cap, int(c[2]) + 64)
        else:
            raise _lib.UwcvError(_lib.E_CAPACITY, "uwcv_union_measure")
        k = int(c[0])
        ui, uf = u_i[:k].cpu().numpy(), u_f[:k].cpu().numpy()
    # reference order: cv2 returns contours in reverse raster order of their start pixel and
    # imutils sorts them (stable) by boundingRect x; rows of different images stay image-major
    order = np.lexsort((-ui[:, 2], -ui[:, 3]))              # (start_y, start_x) descending
    ui, uf = ui[order], uf[order]
    order = np.lexsort((ui[:, 4], ui[:, 0]))                # stable: image, then brect_x
    return UnionTable(ui[order], uf[order])
