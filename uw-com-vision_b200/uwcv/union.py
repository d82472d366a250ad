"""Union / connected-component mode: the reference-literal ``GetMask_Contours``.

nn_inference.py:394-459 ORs the masks of the class of interest into ONE image and measures
every external contour of that union (contourArea >= 100), left to right: touching instances
merge and an instance with several blobs yields several rows.  ``measure_union`` reproduces
those rows; everything between the predictor output and the row table runs in libuwcv.so --
paste, grouping of the instances by box overlap (``uwcv_union_group``), OR, border following,
descriptors -- with ONE synchronising read at the end; the host orders the small row table.
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
        if not bool(keep.all()):                       # (dropping nothing: no copy of the masks)
            b, masks = b[keep], masks[keep]
        bl.append(b.cpu())
        ml.append(masks.to(torch.float32).reshape(-1, MASK_SIDE, MASK_SIDE).cpu())
        il.append(torch.full((int(b.shape[0]),), image_idx_offset + k, dtype=torch.int32))
    boxes = torch.cat(bl)
    n = int(boxes.shape[0])
    if n == 0:
        return empty
    L = eng.L
    B = len(batch)
    st = _stream_ptr(dev)
    with torch.cuda.device(dev):
        d_boxes = boxes.contiguous().to(dev)
        # the masks are concatenated straight into pinned memory (one host copy) and DMA'd from there
        slot = eng.slot(0)
        if slot.pending is not None:
            slot.pending.result()
        h_masks = slot.pinned("masks", (n, MASK_SIDE, MASK_SIDE), torch.float32)
        torch.cat(ml, out=h_masks)
        d_masks = slot.device("masks", (n, MASK_SIDE, MASK_SIDE), torch.float32)
        d_masks.copy_(h_masks, non_blocking=True)
        d_img = torch.cat(il).to(dev)
        rows_i = torch.empty((n, NUM_INT), dtype=torch.int64, device=dev)
        rows_f = torch.empty((n, NUM_FLOAT), dtype=torch.float64, device=dev)
        # layout + paste (cropped: no full-frame planes), no per-instance contour pass
        member_words = tile_words(boxes, H, W)
        eng.run(d_masks, d_boxes, H, W, image_idx=d_img, threshold=mask_threshold,
                pixels_per_metric=pixels_per_metric, rows_i=rows_i, rows_f=rows_f, stages=3,
                n_tile_words=member_words)
        ws = eng._ws
        # groups on the device (no host round trip between the paste and the contour kernels)
        gws = torch.empty(L.uwcv_union_group_workspace_bytes(n, B), dtype=torch.uint8, device=dev)
        d_member = torch.empty(n, dtype=torch.int32, device=dev)
        d_gdesc = torch.empty(n * 32, dtype=torch.uint8, device=dev)
        d_gimg = torch.empty(n, dtype=torch.int32, device=dev)
        gcount = torch.zeros(4, dtype=torch.int64, device=dev)
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
        # capacities: a group tile can exceed the sum of its members' tiles (a diagonal chain of
        # boxes), a contour count / extreme-row count is data dependent: generous first guesses,
        # exact sizes reported by the device when one of them is too small
        gwords_cap = 4 * int(member_words) + 4096
        rec_cap = 2 * n + 256
        ext_cap = 8 * int(boxes[:, 3].sub(boxes[:, 1]).clamp(min=0).sum().item()) + 16 * n + 1024
        for _attempt in range(3):
            gwords_cap = (gwords_cap + 3) & ~3
            gplanes = torch.empty(3 * gwords_cap, dtype=torch.int32, device=dev)
            rec_ws = torch.empty(L.uwcv_union_workspace_bytes(rec_cap, ext_cap), dtype=torch.uint8,
                                 device=dev)
            u_i = torch.zeros((rec_cap, 10), dtype=torch.int64, device=dev)
            u_f = torch.zeros((rec_cap, 16), dtype=torch.float64, device=dev)
            _lib.check(L.uwcv_union_group(_ptr(rows_i), n, B, int(image_idx_offset), _ptr(gws), gws.numel(),
                                          _ptr(d_member), _ptr(d_gdesc), _ptr(d_gimg), gwords_cap,
                                          _ptr(gcount), st), "uwcv_union_group")
            _lib.check(L.uwcv_union_measure_grouped(_ptr(ws), ws.numel(), n, _ptr(d_member), _ptr(d_gdesc),
                                                    _ptr(d_gimg), _ptr(gcount), _ptr(gplanes), gwords_cap,
                                                    _ptr(rec_ws), rec_ws.numel(), rec_cap, ext_cap,
                                                    float(pixels_per_metric), _ptr(u_i), _ptr(u_f),
                                                    _ptr(counters), st), "uwcv_union_measure_grouped")
            eng.launches += 6
            # ONE synchronising read for the whole call: status, counters and the rows together
            pin = torch.empty(12, dtype=torch.int64).pin_memory()
            pin[0:4].copy_(eng.status, non_blocking=True)
            pin[4:8].copy_(gcount, non_blocking=True)
            pin[8:12].copy_(counters, non_blocking=True)
            h_i = torch.empty((rec_cap, 10), dtype=torch.int64).pin_memory()
            h_f = torch.empty((rec_cap, 16), dtype=torch.float64).pin_memory()
            h_i.copy_(u_i, non_blocking=True)
            h_f.copy_(u_f, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            c = pin.tolist()
            if c[0] != 0:
                raise _lib.UwcvError(int(c[0]), f"uwcv_paste_measure (needs {int(c[1])} tile words)")
            if c[7] == 0 and c[9] == 0:
                break
            gwords_cap = max(gwords_cap, int(c[5]) + 64)     # exact needs reported by the device
            rec_cap = max(rec_cap, int(c[8]) + 64)
            ext_cap = max(ext_cap, int(c[10]) + 64)
        else:
            raise _lib.UwcvError(_lib.E_CAPACITY, "uwcv_union_measure")
        k = int(c[8])
        if k == 0:
            return empty
        ui, uf = h_i[:k].numpy().copy(), h_f[:k].numpy().copy()
    # reference order: cv2 returns contours in reverse raster order of their start pixel and
    # imutils sorts them (stable) by boundingRect x; rows of different images stay image-major
    order = np.lexsort((-ui[:, 2], -ui[:, 3]))              # (start_y, start_x) descending
    ui, uf = ui[order], uf[order]
    order = np.lexsort((ui[:, 4], ui[:, 0]))                # stable: image, then brect_x
    return UnionTable(ui[order], uf[order])
