"""Host-side mirror of the reference's call surface for the hot path.

Same names, argument meaning and error behaviour as the Detectron2 functions the
reference reaches through ``predictor(im)`` (nn_inference.py:372) and as its own
measurement entry ``GetMask_Contours`` (nn_inference.py:371-459):

* ``paste_masks_in_image(masks, boxes, image_shape, threshold)``  -> N x H x W bool
* ``detector_postprocess(results, output_height, output_width, mask_threshold)``
* ``fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh,
  nms_thresh, topk_per_image)``
* ``measure_instances(instances, output_size, classes_of_interest, ...)`` -> table

PyTorch is used for device memory, streams and the three N x 4 box ops that must stay
bit-identical to Detectron2's (scale, clip, non-empty); every pixel- or box-pair-level
computation runs in libuwcv.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .schema import (CSV_SOURCE, FCOL, FLOAT_COLUMNS, ICOL, INT_COLUMNS, NUM_FLOAT, NUM_INT)
from .structures import Boxes, Instances

MASK_SIDE = 28


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("uwcv needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"uwcv runs on CUDA devices only, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ----------------------------------------------------------------------------------
# box glue: the three Detectron2 ops that must be bit-identical
# ----------------------------------------------------------------------------------

def scale_clip_boxes(boxes: torch.Tensor, in_size: Tuple[int, int], out_size: Tuple[int, int],
                     check: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Boxes.scale + Boxes.clip + Boxes.nonempty of detector_postprocess.
    Returns (output-space boxes, keep mask); works on whatever device ``boxes`` is on.
    Same float32 operations as Detectron2's column-wise ``tensor[:, 0::2] *= scale_x`` /
    ``clamp(min=0, max=w)`` (one IEEE multiply by the float32-rounded scale, one min/max per
    element), written as three whole-tensor ops: the strided column form costs 3.5 ms per
    64 000 boxes on one host thread, this one 0.2 ms."""
    out_h, out_w = int(out_size[0]), int(out_size[1])
    scale_x, scale_y = out_w / in_size[1], out_h / in_size[0]
    kw = dict(dtype=torch.float32, device=boxes.device)
    b = boxes.to(torch.float32)
    if scale_x == scale_y:                  # (broadcast forms are several times slower on the host)
        b = b * scale_x if scale_x != 1.0 else b.clone()
    else:
        b = b * torch.tensor([scale_x, scale_y, scale_x, scale_y], **kw)
    # (``check=False``: the caller verifies finiteness itself, without a host synchronisation)
    if check and b.numel() and not bool(b.abs().max() < float("inf")):          # inf or NaN anywhere
        raise AssertionError("Box tensor contains infinite or NaN!")   # Boxes.clip asserts
    if out_w == out_h:
        b = b.clamp_(min=0, max=out_w)
    else:
        b = torch.clamp(b, min=torch.zeros(4, **kw),
                        max=torch.tensor([out_w, out_h, out_w, out_h], **kw))
    size = b[:, 2:] - b[:, :2]
    keep = (size[:, 0] > 0) & (size[:, 1] > 0)
    return b, keep


def tile_words(boxes: torch.Tensor, H: int, W: int) -> int:
    """Exact number of 32-pixel tile words the layout kernel will allocate for these
    output-space boxes (host mirror of csrc/paste_measure.cu::tile_geometry)."""
    if boxes.numel() == 0:
        return 0
    return int(tile_words_each(boxes, H, W).sum().item())


def tile_words_each(boxes: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Per-instance tile words (int64 tensor on the device of ``boxes``)."""
    b = boxes.detach().to(torch.float32)
    lo, hi = b[:, :2], b[:, 2:]
    size = hi - lo                                         # float32, as the kernel computes it
    ok = (size > 0).all(dim=1) & torch.isfinite(b).all(dim=1)
    m = size.double() / MASK_SIDE + 1.0
    fa, fb = torch.floor(lo.double() - m), torch.ceil(hi.double() + m)
    lim = torch.tensor([W - 1, H - 1], dtype=torch.float64, device=b.device)
    ok &= ~((fb < 0) | (fa > lim)).any(dim=1)
    pa = torch.minimum(fa.clamp(min=0), lim).nan_to_num(0).long()
    pb = torch.minimum(fb.clamp(min=0), lim).nan_to_num(0).long()
    tw = (pb[:, 0] >> 5) - (pa[:, 0] >> 5) + 1
    th = pb[:, 1] - pa[:, 1] + 1
    return torch.where(ok, tw * th, torch.zeros_like(tw))


def tile_words_bound(boxes: torch.Tensor) -> int:
    """Cheap upper bound of ``tile_words`` (never below it): the window of an instance spans at
    most ``size * 30 / 28 + 4`` pixels per axis before clipping, i.e. at most that / 32 + 2
    words and that + 1 rows.  Used to size the workspace when a call must not be repeated
    (gathered calls: a retry on one rank only would desynchronise the ranks)."""
    if boxes.numel() == 0:
        return 0
    b = boxes.detach().to(torch.float32)
    span = ((b[:, 2:] - b[:, :2]).clamp(min=0) * (30.0 / 28.0) + 5.0).nan_to_num(0.0, posinf=0.0)
    words = (torch.floor(span[:, 0] / 32.0) + 3.0) * (span[:, 1] + 2.0)
    return int(words.sum(dtype=torch.float64).item()) + 1


# ----------------------------------------------------------------------------------
# result table
# ----------------------------------------------------------------------------------

@dataclass
class MeasurementTable:
    """int64 [R, 20] + float64 [R, 30] measurement rows (schema.py)."""
    ints: np.ndarray
    floats: np.ndarray

    INT_COLUMNS = INT_COLUMNS
    FLOAT_COLUMNS = FLOAT_COLUMNS

    def __len__(self) -> int:
        return self.ints.shape[0]

    def __getitem__(self, name: str) -> np.ndarray:
        if name in ICOL:
            return self.ints[:, ICOL[name]]
        if name in FCOL:
            return self.floats[:, FCOL[name]]
        raise KeyError(name)

    def select(self, mask) -> "MeasurementTable":
        return MeasurementTable(self.ints[mask], self.floats[mask])

    def for_class(self, class_id: int) -> "MeasurementTable":
        return self.select(self["class_id"] == class_id)

    def reference_rows(self, min_contour_area: float = 100.0) -> np.ndarray:
        """K x 9 rows in the reference's CSV column order (nn_inference.py:561, :569) for the
        instances that pass its ``contourArea < 100 -> skip`` cut (nn_inference.py:412)."""
        ok = (self["valid"] == 1) & (self["contour_area"] >= min_contour_area)
        cols = [FCOL[c] for c in CSV_SOURCE]
        return self.floats[ok][:, cols]

    @staticmethod
    def empty() -> "MeasurementTable":
        return MeasurementTable(np.zeros((0, NUM_INT), np.int64), np.zeros((0, NUM_FLOAT), np.float64))

    @staticmethod
    def concat(tables: Sequence["MeasurementTable"]) -> "MeasurementTable":
        if not tables:
            return MeasurementTable.empty()
        return MeasurementTable(np.concatenate([t.ints for t in tables], axis=0),
                                np.concatenate([t.floats for t in tables], axis=0))

    def to_dataframe(self):
        import pandas as pd
        df = pd.DataFrame(self.ints, columns=list(INT_COLUMNS))
        for j, c in enumerate(FLOAT_COLUMNS):
            df[c] = self.floats[:, j]
        return df


class HostLease:
    """Owner object of a span of (pinned or shared) host memory handed to the caller as numpy
    arrays: ``array()`` is an int64 view whose ``.base`` is the lease, every view derived from
    it keeps the lease alive, and ``on_release`` runs once when the last one is gone."""

    def __init__(self, ptr: int, n_int64: int, on_release, keep=None):
        import weakref
        self.__array_interface__ = {"shape": (int(n_int64),), "typestr": "<i8",
                                    "data": (int(ptr), False), "version": 3}
        self._keep = keep                      # whatever owns the memory (tensor, mmap)
        weakref.finalize(self, on_release)

    def array(self) -> np.ndarray:
        return np.asarray(self)


# ----------------------------------------------------------------------------------
# the engine: buffers + the C-ABI calls
# ----------------------------------------------------------------------------------

class _CompressibleBuffer:
    """A uwcv_planes_alloc allocation exposed through ``__cuda_array_interface__`` (torch.as_tensor
    shares it and keeps this object alive for as long as any tensor views the memory)."""

    def __init__(self, eng, ptr: int, words: int, compressed: bool):
        self._eng, self._ptr, self.compressed = eng, ptr, compressed
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i4", "data": (ptr, False),
                                         "version": 3, "strides": None}

    @staticmethod
    def create(eng, words: int):
        ptr, comp = C.c_void_p(), C.c_int(0)
        with torch.cuda.device(eng.device):
            rc = eng.L.uwcv_planes_alloc(int(words) * 4, C.byref(ptr), C.byref(comp))
        if rc != 0 or not ptr.value:
            return None
        if not comp.value:                       # ordinary pages: torch's allocator serves those better
            eng.L.uwcv_planes_free(ptr)
            return None
        return _CompressibleBuffer(eng, int(ptr.value), words, True)

    def __del__(self):
        ptr, self._ptr = getattr(self, "_ptr", None), None
        if ptr:
            try:
                torch.cuda.synchronize(self._eng.device)     # like cudaFree: nothing may still use it
                self._eng.L.uwcv_planes_free(C.c_void_p(ptr))
            except Exception:                                # interpreter shutdown
                pass


class Engine:
    """Per-device cache of workspace / output buffers around the C ABI.

    ``run`` enqueues layout + paste/measure + contour kernels on the current stream and
    returns device tensors; nothing synchronises until the caller reads them.  One engine per
    device, driven from one host thread at a time (the C ABI underneath is re-entrant; the
    buffer caches here are not locked)."""

    _engines = {}
    _engines_lock = threading.Lock()

    @classmethod
    def get(cls, device=None) -> "Engine":
        device = _require_cuda(device)
        with cls._engines_lock:                 # two threads asking at once get the same engine
            e = cls._engines.get(device)
            if e is None:
                e = cls._engines[device] = Engine(device)
        return e

    def __init__(self, device: torch.device):
        self.device = device
        self.L = _lib.lib()
        self._ws: Optional[torch.Tensor] = None
        self._ws2: Optional[torch.Tensor] = None
        self._cap_words = 0
        self._cap_n = 0
        self.status = torch.zeros(4, dtype=torch.int64, device=device)
        self.launches = 0            # kernels enqueued by this engine (for bench accounting)
        self.h2d_stream = torch.cuda.Stream(device)
        self.small_stream = torch.cuda.Stream(device)
        self.d2h_stream = torch.cuda.Stream(device)
        # the border trace of call i runs on its own stream so that its (latency-bound, mostly
        # idle) tail overlaps the (HBM-bound) paste of call i + 1: two workspaces, used in turn
        self.trace_stream = torch.cuda.Stream(device, priority=-1)   # its CTAs go first when SMs free up
        self._trace_done = [None, None]
        # split pipeline: the planes are written from the tiles by a data-movement kernel on its
        # own stream (csrc/plane_fill.cu) beside the trace of the same call and the tile kernel of
        # the next one; its CTAs are placed first when an SM has room
        self.fill_stream = torch.cuda.Stream(device, priority=-2)
        self._fill_done = [None, None]
        self.planes_done = None      # event behind the last plane fill (split pipeline)
        self.fill_events = None      # list: (start, end) timing events of every plane fill are appended
        self._parity = 0
        self.split_default = True
        self.compressible_planes = True     # plane buffers in compressible device memory when granted
        self._slots = {}
        self._host_pool = []
        self._fused = None

    def host_rows(self, r: int):
        """Pinned host memory for one call's rows: returns (torch int64 [r,20], torch float64
        [r,30], and the numpy views of the same memory that the MeasurementTable will own).
        Ownership is explicit: the numpy views hang off a ``HostLease`` (their ``.base``), and
        the buffer goes back to the pool when the lease dies, i.e. when the table AND every
        slice the caller took from it are gone -- no reference counting by hand."""
        need = max(r * (NUM_INT + NUM_FLOAT), 1)
        entry = None
        for e in self._host_pool:
            if e["free"] and e["buf"].numel() >= need:
                entry = e
                break
        if entry is None:
            # drop free buffers that are too small; leased ones stay until their lease ends
            self._host_pool = [e for e in self._host_pool if not e["free"]]
            buf = torch.empty(int(need * 1.1) + 64, dtype=torch.int64).pin_memory()
            entry = {"buf": buf, "free": True}
            self._host_pool.append(entry)
        entry["free"] = False
        buf = entry["buf"]

        def give_back(e=entry):
            e["free"] = True

        arr = HostLease(buf.data_ptr(), buf.numel(), give_back, keep=buf).array()
        a, b = r * NUM_INT, r * (NUM_INT + NUM_FLOAT)
        t_i = buf[:a].view(r, NUM_INT)
        t_f = buf[a:b].view(torch.float64).view(r, NUM_FLOAT)
        n_i = arr[:a].reshape(r, NUM_INT)
        n_f = arr[a:b].view(np.float64).reshape(r, NUM_FLOAT)
        return t_i, t_f, n_i, n_f

    def fused_gather(self):
        """The symmetric-memory gather context of this device (created on first use, a
        collective), or None when torch.distributed is not a multi-rank NCCL job or symmetric
        memory cannot be set up (the caller then uses the NCCL all_gather_table)."""
        if self._fused is None:
            self._fused = False
            import os
            import torch.distributed as dist
            if not os.environ.get("UWCV_NO_FUSED_GATHER") and dist.is_available() and \
                    dist.is_initialized() and dist.get_world_size() > 1 and \
                    dist.get_backend() == "nccl":
                ok = 1
                try:
                    from .dist import FusedGather
                    fg = FusedGather(self.device)
                    fg.ensure(1024)
                except Exception as e:                      # noqa: BLE001
                    import warnings
                    warnings.warn(f"uwcv: fused gather unavailable ({type(e).__name__}: {e}); "
                                  "using the NCCL all-gather")
                    ok = 0
                # all ranks take the same path
                t = torch.tensor([ok], dtype=torch.int32, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if int(t.item()) == 1:
                    self._fused = fg
        return self._fused or None

    def host_table(self):
        """The node-shared host table of this device's process group (created on first use, a
        collective), or None when it cannot be set up on every rank (not one node, shared memory
        or cudaHostRegister unavailable): the caller then gathers on the devices."""
        if getattr(self, "_host_table", None) is None:
            self._host_table = False
            import os
            import torch.distributed as dist
            if dist_is_multi() and dist.get_backend() == "nccl":
                one_node = int(os.environ.get("LOCAL_WORLD_SIZE", "0")) == dist.get_world_size()
                ok = 1 if one_node else 0
                ht = None
                if ok:
                    try:
                        from .dist import SharedHostTable
                        ht = SharedHostTable(self.device)
                        ht.ensure(1024)
                    except Exception as e:                  # noqa: BLE001
                        import warnings
                        warnings.warn(f"uwcv: shared host table unavailable ({type(e).__name__}: {e}); "
                                      "gathering on the devices")
                        ok = 0
                t = torch.tensor([ok], dtype=torch.int32, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if int(t.item()) == 1:
                    self._host_table = ht
        return self._host_table or None

    def slot(self, k: int = 0) -> "_Slot":
        s = self._slots.get(k)
        if s is None:
            s = self._slots[k] = _Slot(self)
        return s

    def _workspace(self, n: int, words: int, slot: int = 0) -> torch.Tensor:
        """Workspace `slot` (0: default; 1: a second buffer of the same capacity for callers
        that keep two calls in flight)."""
        if self._ws is None or n > self._cap_n or words > self._cap_words:
            if self._ws is not None:
                torch.cuda.synchronize(self.device)     # another stream may still trace on it
            cap_n = max(n, self._cap_n)
            cap_w = max(int(words * 1.25) + 1024, self._cap_words)
            nbytes = self.L.uwcv_workspace_bytes(cap_n, cap_w)
            self._ws = None
            self._ws2 = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._cap_n, self._cap_words = cap_n, cap_w
        if slot == 0:
            return self._ws
        if self._ws2 is None or self._ws2.numel() != self._ws.numel():
            self._ws2 = torch.empty(self._ws.numel(), dtype=torch.uint8, device=self.device)
        return self._ws2

    def run(self, masks: torch.Tensor, boxes: torch.Tensor, H: int, W: int, *,
            image_idx: Optional[torch.Tensor] = None, inst_idx: Optional[torch.Tensor] = None,
            classes: Optional[torch.Tensor] = None, scores: Optional[torch.Tensor] = None,
            threshold: float = 0.5, pixels_per_metric: float = 0.85,
            planes: Optional[torch.Tensor] = None, n_tile_words: Optional[int] = None,
            rows_i: Optional[torch.Tensor] = None, rows_f: Optional[torch.Tensor] = None,
            stages: int = 7, status: Optional[torch.Tensor] = None, ws_slot: int = 0,
            first: int = 0, count: Optional[int] = None, mask_channels: int = 1,
            channel_offset: int = 0, logits: bool = False, gather=None):
        """All tensors on ``self.device``, contiguous: masks [N,28,28] f32, boxes [N,4] f32
        (output space), image_idx/inst_idx int32, classes int64, scores f32,
        planes None or uint32/int32 [N, H, plane_row_words(W)].  Single-forward form: masks
        [N, mask_channels, 28, 28] (instance i uses channel classes[i] + channel_offset),
        ``logits=True`` applies the mask head's sigmoid while staging (uwcv_paste_measure_heads).
        Returns (rows_i [N,20] int64, rows_f [N,30] float64, status [4] int64)."""
        n = int(boxes.shape[0])
        dev = self.device
        if rows_i is None:
            rows_i = torch.empty((n, NUM_INT), dtype=torch.int64, device=dev)
        if rows_f is None:
            rows_f = torch.empty((n, NUM_FLOAT), dtype=torch.float64, device=dev)
        if n_tile_words is None:
            n_tile_words = tile_words(boxes, H, W)
        ws = self._workspace(n, n_tile_words, ws_slot)
        status = self.status if status is None else status
        if (stages & 1) and self._trace_done[ws_slot] is not None:
            # the layout is about to hand this workspace out again: after the trace that reads it
            torch.cuda.current_stream(dev).wait_event(self._trace_done[ws_slot])
            self._trace_done[ws_slot] = None
        if (stages & 1) and self._fill_done[ws_slot] is not None:
            torch.cuda.current_stream(dev).wait_event(self._fill_done[ws_slot])   # ... and the plane fill
            self._fill_done[ws_slot] = None
        with torch.cuda.device(dev):
            rc = self.L.uwcv_paste_measure_gather(
                _ptr(masks), int(mask_channels), int(channel_offset), int(bool(logits)), _ptr(boxes), _ptr(image_idx), _ptr(inst_idx), _ptr(classes),
                _ptr(scores), n, int(H), int(W), float(threshold), float(pixels_per_metric),
                _ptr(planes), _ptr(rows_i), _ptr(rows_f), _ptr(ws), ws.numel(),
                _ptr(status), _stream_ptr(dev), int(stages), int(first),
                int(n - first if count is None else count),
                C.byref(gather) if (gather is not None and (stages & 4)) else None)
        _lib.check(rc, "uwcv_paste_measure")
        if n > 0:        # layout = 3 kernels, paste = 1 (rows only / split: 2), contour = 1, plane fill = 1
            self.launches += 3 * (stages & 1) \
                + ((stages >> 1) & 1) * (1 if (planes is not None and not (stages & 16)) else 2) \
                + ((stages >> 2) & 1) + ((stages >> 3) & 1)
        return rows_i, rows_f, status

    def run_overlapped(self, masks, boxes, H, W, *, paste_ranges=None, after=None, split=None,
                       fill_once=False, **kw):
        """Layout + paste on the current stream, border trace on ``trace_stream``, alternating
        between the two workspaces.  ``paste_ranges``: [(first, count, event or None), ...] --
        the paste of a range waits for its event (chunks of a host->device copy in flight).
        ``after``: callable run on the trace stream behind the trace (collectives, D2H hand-off).
        ``split`` (default: whenever planes are written): the paste writes tiles and integer rows
        only (compute-bound) and the planes are written from the tiles on ``fill_stream`` by a
        kernel that only moves data, so the HBM-bound fill shares the SMs with the trace of this
        call and the tile kernel of the next; ``self.planes_done`` is the event behind the fill.
        ``fill_once``: ONE plane fill over the whole call behind the last range's tiles instead of
        one per range (a fill launch pays its ramp and its drain -- ~0.1 ms -- whatever its size;
        callers that keep three or more calls in flight have the previous call's fill to run under
        the ranges still arriving, so nothing is gained by starting this call's fill early).
        Returns the event that marks the rows complete."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        p = self._parity
        self._parity ^= 1
        n = int(boxes.shape[0])
        kw = dict(kw, ws_slot=p)
        if split is None:
            split = self.split_default
        split = bool(split) and kw.get("planes") is not None
        sbit = 16 if split else 0
        self.run(masks, boxes, H, W, stages=1, **kw)
        def fill(first, count):
            tiles = torch.cuda.Event()
            tiles.record(main)
            with torch.cuda.stream(self.fill_stream):
                self.fill_stream.wait_event(tiles)
                if self.fill_events is not None:       # (bench: live time of the fill inside the pipeline)
                    t0 = torch.cuda.Event(enable_timing=True)
                    t0.record(self.fill_stream)
                self.run(masks, boxes, H, W, stages=8 | sbit, first=first, count=count, **kw)
                if self.fill_events is not None:
                    t1 = torch.cuda.Event(enable_timing=True)
                    t1.record(self.fill_stream)
                    self.fill_events.append((t0, t1))

        ranges = list(paste_ranges or [(0, n, None)])
        fill_once = bool(fill_once) and split and len(ranges) > 1
        for first, count, ev in ranges:
            if ev is not None:
                main.wait_event(ev)
            if count > 0:
                self.run(masks, boxes, H, W, stages=2 | sbit, first=first, count=count, **kw)
                if split and not fill_once:
                    fill(first, count)
        if fill_once:
            lo = min(f for f, c, _ in ranges)
            hi = max(f + c for f, c, _ in ranges)
            if hi > lo:
                fill(lo, hi - lo)
        pasted = torch.cuda.Event()
        pasted.record(main)
        if split:
            fd = torch.cuda.Event()
            fd.record(self.fill_stream)
            self._fill_done[p] = fd
            self.planes_done = fd
        with torch.cuda.stream(self.trace_stream):
            self.trace_stream.wait_event(pasted)
            self.run(masks, boxes, H, W, stages=4 | sbit, **kw)
            if after is not None:
                after()
            done = torch.cuda.Event()
            done.record(self.trace_stream)
        self._trace_done[p] = done
        return done

    def check_status(self) -> None:
        """Synchronising read of the device status word; raises on workspace overflow."""
        st = self.status.cpu()
        if int(st[0]) != 0:
            raise _lib.UwcvError(int(st[0]), f"uwcv_paste_measure (needs {int(st[1])} tile words)")

    def plane_buffer(self, words: int) -> torch.Tensor:
        """int32 device tensor of ``words`` elements for full-frame bit-planes, in COMPRESSIBLE
        device memory when the driver grants it (uwcv_planes_alloc: the planes are zeros almost
        everywhere and B200 compresses such pages between L2 and HBM -- the plane fill runs 14 %
        faster, a read-back 40 %), else an ordinary torch allocation.  The memory is returned to
        the driver when the last tensor viewing it dies (uwcv_planes_free behind a device
        synchronisation, as cudaFree does)."""
        words = max(int(words), 1)
        if self.compressible_planes:
            buf = _CompressibleBuffer.create(self, words)
            if buf is not None:
                with torch.cuda.device(self.device):
                    return torch.as_tensor(buf, device=self.device)
            self.compressible_planes = False          # not granted on this device: do not ask again
        return torch.empty(words, dtype=torch.int32, device=self.device)

    def scratch_planes(self, n: int, H: int, W: int) -> torch.Tensor:
        """Engine-owned plane buffer for callers that want the Detectron2-literal masks
        written to HBM but do not take ownership (reused by the next call)."""
        wpr = self.L.uwcv_plane_row_words(int(W))
        need = n * H * wpr
        buf = getattr(self, "_planes", None)
        if buf is None or buf.numel() < need:
            if buf is not None:
                self.fill_stream.synchronize()        # a plane fill may still be writing the old buffer
            self._planes = None
            self._planes = buf = self.plane_buffer(need)
        return buf[:need].view(n, H, wpr)

    def alloc_planes(self, n: int, H: int, W: int) -> torch.Tensor:
        wpr = self.L.uwcv_plane_row_words(int(W))
        return self.plane_buffer(n * H * wpr)[:n * H * wpr].view(n, H, wpr)

    def unpack(self, planes: torch.Tensor, H: int, W: int) -> torch.Tensor:
        n = int(planes.shape[0])
        out = torch.empty((n, H, W), dtype=torch.bool, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.L.uwcv_unpack_planes(_ptr(planes), n, int(H), int(W), _ptr(out),
                                           _stream_ptr(self.device))
        _lib.check(rc, "uwcv_unpack_planes")
        self.launches += 1 if n > 0 else 0
        return out

    def nms(self, boxes: torch.Tensor, scores: torch.Tensor, classes: torch.Tensor,
            image_off: Sequence[int], score_thresh: float, nms_thresh: float, topk: int,
            num_classes: int = 0):
        """boxes [R,4] f32, scores [R] f32, classes [R] i64 on device; image_off host ints [B+1];
        num_classes = K of the box head (0: 128).
        Returns (keep [R] int64, keep_count [B] int32) device tensors."""
        B = len(image_off) - 1
        off = (C.c_int64 * (B + 1))(*[int(v) for v in image_off])
        R = int(image_off[-1])
        dev = self.device
        keep = torch.empty(max(R, 1), dtype=torch.int64, device=dev)
        cnt = torch.zeros(max(B, 1), dtype=torch.int32, device=dev)
        nbytes = self.L.uwcv_nms_workspace_bytes(off, B, int(num_classes))
        ws = getattr(self, "_nms_ws", None)           # O(R) bytes, cached across calls
        if ws is None or ws.numel() < nbytes:
            ws = self._nms_ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = self.L.uwcv_nms_filter(_ptr(boxes), _ptr(scores), _ptr(classes), off, B,
                                        int(num_classes), float(score_thresh), float(nms_thresh), int(topk),
                                        _ptr(keep), _ptr(cnt), _ptr(ws), ws.numel(),
                                        _stream_ptr(dev))
        _lib.check(rc, "uwcv_nms_filter")
        self.launches += (B + 32) // 32 + (4 if R > 0 else 2)    # offsets, buckets, segment sorts, sweep, merge
        return keep, cnt


# ----------------------------------------------------------------------------------
# Detectron2-shaped entry points
# ----------------------------------------------------------------------------------

def _as_box_tensor(boxes) -> torch.Tensor:
    return boxes.tensor if hasattr(boxes, "tensor") else boxes


def paste_masks_in_image(masks: torch.Tensor, boxes, image_shape: Tuple[int, int],
                         threshold: float = 0.5, *, packed: bool = False,
                         device=None) -> torch.Tensor:
    """detectron2.layers.mask_ops.paste_masks_in_image: N x Hm x Wm probabilities + N x 4
    boxes (already in image coordinates) -> N x H x W bool on the CUDA device.
    ``packed=True`` returns the bit-planes [N, H, plane_row_words(W)] int32 instead."""
    if masks.shape[-1] != masks.shape[-2]:
        raise AssertionError("Only square mask predictions are supported")
    if masks.shape[-1] != MASK_SIDE and len(masks) > 0:
        raise ValueError(f"uwcv is built for {MASK_SIDE}x{MASK_SIDE} mask heads, got {tuple(masks.shape)}")
    if not threshold > 0:
        raise ValueError("uwcv requires a positive mask threshold (Detectron2's soft-mask "
                         "mode threshold < 0 is not part of the reference path)")
    boxes = _as_box_tensor(boxes)
    if len(boxes) != len(masks):
        raise AssertionError(boxes.shape)
    H, W = int(image_shape[0]), int(image_shape[1])
    dev = _require_cuda(device if device is not None else
                        (boxes.device if boxes.is_cuda else None))
    eng = Engine.get(dev)
    n = len(masks)
    if n == 0:
        if packed:
            return eng.alloc_planes(0, H, W)
        return torch.zeros((0, H, W), dtype=torch.bool, device=dev)
    m = masks.reshape(n, MASK_SIDE, MASK_SIDE).to(dev, torch.float32).contiguous()
    b = boxes.to(dev, torch.float32).contiguous()
    planes = eng.alloc_planes(n, H, W)
    eng.run(m, b, H, W, threshold=threshold, planes=planes)
    eng.check_status()
    return planes if packed else eng.unpack(planes, H, W)


def detector_postprocess(results, output_height: int, output_width: int,
                         mask_threshold: float = 0.5) -> Instances:
    """detectron2.modeling.postprocessing.detector_postprocess: rescale boxes to the output
    size, clip, drop empty ones, paste ``pred_masks`` (N x 1 x 28 x 28) to N x H x W bool."""
    fields = dict(results.get_fields()) if hasattr(results, "get_fields") else dict(results._fields)
    box_key = "pred_boxes" if "pred_boxes" in fields else "proposal_boxes"
    if box_key not in fields:
        raise AssertionError("Predictions must contain boxes!")
    b, keep = scale_clip_boxes(_as_box_tensor(fields[box_key]), results.image_size,
                               (output_height, output_width))
    out = Instances((output_height, output_width))
    for k, v in fields.items():
        if k == box_key:
            out.set(k, Boxes(b[keep]))
        else:
            out.set(k, v[keep])
    if out.has("pred_masks"):
        pm = out.pred_masks
        out.remove("pred_masks")
        out.set("pred_masks", paste_masks_in_image(pm[:, 0, :, :], out.get(box_key).tensor,
                                                   (output_height, output_width),
                                                   threshold=mask_threshold))
    return out


def fast_rcnn_inference_single_image(boxes: torch.Tensor, scores: torch.Tensor,
                                     image_shape: Tuple[int, int], score_thresh: float,
                                     nms_thresh: float, topk_per_image: int, device=None):
    """detectron2 fast_rcnn_inference_single_image: boxes R x (K*4) (or R x 4), scores
    R x (K+1) with the background column last.  Returns (Instances, kept proposal rows) -- the
    rows index the input after non-finite rows have been dropped, as Detectron2's do."""
    dev = _require_cuda(device if device is not None else (boxes.device if boxes.is_cuda else None))
    boxes = boxes.to(dev, torch.float32)
    scores = scores.to(dev, torch.float32)
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not bool(valid.all()):
        boxes, scores = boxes[valid], scores[valid]
    scores = scores[:, :-1]
    R, K = scores.shape
    nreg = boxes.shape[1] // 4
    h, w = image_shape
    bx = boxes.reshape(-1, 4)
    bx = torch.stack((bx[:, 0].clamp(min=0, max=w), bx[:, 1].clamp(min=0, max=h),
                      bx[:, 2].clamp(min=0, max=w), bx[:, 3].clamp(min=0, max=h)), dim=-1)
    bx = bx.view(R, nreg, 4)
    if nreg == 1:
        bx = bx.expand(R, K, 4)
    cand_boxes = bx.reshape(R * K, 4).contiguous()           # candidate (r, k) at row r*K + k
    cand_scores = scores.reshape(R * K).contiguous()
    cand_cls = torch.arange(K, device=dev, dtype=torch.int64).repeat(R)
    eng = Engine.get(dev)
    keep, cnt = eng.nms(cand_boxes, cand_scores, cand_cls, [0, R * K], score_thresh, nms_thresh,
                        topk_per_image, num_classes=K)
    k = int(cnt[0].item())
    keep = keep[:k]
    res = Instances(tuple(image_shape))
    res.pred_boxes = Boxes(cand_boxes[keep])
    res.scores = cand_scores[keep]
    res.pred_classes = cand_cls[keep]
    # Detectron2 returns filter_inds[:, 0]: row positions AFTER its finite-value filter
    return res, keep // K


# ----------------------------------------------------------------------------------
# measurement entry (GetMask_Contours / GetCounts replacement)
# ----------------------------------------------------------------------------------

def _has(inst, name: str) -> bool:
    return inst.has(name) if hasattr(inst, "has") else name in inst._fields


def _mask_field(inst):
    """(tensor [N, C, 28, 28], is_logits): ``pred_masks`` (N x 1 x 28 x 28 probabilities, the
    output of mask_rcnn_inference) or, on the single-forward path, ``pred_mask_logits``
    (N x K x 28 x 28 raw mask-head output; channel select + sigmoid happen in the kernel)."""
    has = inst.has if hasattr(inst, "has") else (lambda k: k in inst._fields)
    if has("pred_mask_logits") and not has("pred_masks"):
        m = inst.pred_mask_logits
        if m.dim() != 4:
            raise ValueError(f"pred_mask_logits must be N x K x 28 x 28, got {tuple(m.shape)}")
        return m, True
    m = inst.pred_masks
    if m.dim() == 3:
        m = m[:, None]
    return m, False


def _gather_fields(inst, classes_of_interest):
    boxes = _as_box_tensor(inst.pred_boxes)
    scores = inst.scores
    classes = inst.pred_classes
    masks, _ = _mask_field(inst)
    if classes_of_interest is not None:
        sel = torch.zeros(len(classes), dtype=torch.bool, device=classes.device)
        for c in classes_of_interest:
            sel |= classes == int(c)
        boxes, scores, classes, masks = boxes[sel], scores[sel], classes[sel], masks[sel]
    return boxes, scores, classes, masks


class _Slot:
    """Device + pinned host buffers of one in-flight call.  A synchronous call uses slot 0;
    a ``MeasurementStream`` rotates over ``depth`` slots so that the host->device copies of
    call i + 1 run under the kernels of call i."""

    def __init__(self, eng: "Engine"):
        self.eng = eng
        self.dev = {}
        self.pin = {}
        self.pending: Optional["PendingTable"] = None

    def device(self, name: str, shape, dtype) -> torch.Tensor:
        need = 1
        for v in shape:
            need *= int(v)
        buf = self.dev.get(name)
        if buf is None or buf.dtype != dtype or buf.numel() < need:
            # growing a buffer other streams may still use: drain the device first (rare)
            torch.cuda.synchronize(self.eng.device)
            self.dev[name] = None
            buf = self.dev[name] = torch.empty(int(need * 1.1) + 64, dtype=dtype,
                                               device=self.eng.device)
        return buf[:need].view(*shape)

    def pinned(self, name: str, shape, dtype) -> torch.Tensor:
        need = 1
        for v in shape:
            need *= int(v)
        buf = self.pin.get(name)
        if buf is None or buf.dtype != dtype or buf.numel() < need:
            torch.cuda.synchronize(self.eng.device)
            self.pin[name] = None
            buf = self.pin[name] = torch.empty(int(need * 1.1) + 64, dtype=dtype).pin_memory()
        return buf[:need].view(*shape)


class PendingTable:
    """Handle of an enqueued ``measure_instances`` call: ``result()`` waits for the
    device->host copy of the rows and returns what ``measure_instances`` returns."""

    def __init__(self, slot: Optional[_Slot] = None, retry=None):
        self._slot = slot
        self._retry = retry
        self._done = None
        self._redo = None
        self.hp_ok = None
        self._after_sync = None
        self._value = None
        self.n = 0
        self.planes = None
        self.return_planes = False
        self._planes_done = None
        self.hp_i = self.hp_f = self.hp_s = self.np_i = self.np_f = None

    @staticmethod
    def ready(value) -> "PendingTable":
        p = PendingTable()
        p._value = (value,)
        return p

    def result(self):
        if self._value is not None:
            return self._value[0]
        self._done.synchronize()
        if self._planes_done is not None:              # split pipeline: the fill has its own stream
            self._planes_done.synchronize()
            self._planes_done = None
        if self._after_sync is not None:
            self._after_sync()
            self._after_sync = None
        if self.hp_ok is not None and int(self.hp_ok[0]) == 0:
            # the optimistic device-input path met an empty or non-finite box: general path
            redo = self._redo
            self._release()
            value = redo()
            self._value = (value,)
            return value
        st = self.hp_s
        code = int(st[0])
        if code != 0:
            self._release()
            if code == _lib.E_CAPACITY and self._retry is not None:
                value = self._retry()
                self._value = (value,)
                return value
            raise _lib.UwcvError(code, f"uwcv_paste_measure (needs {int(st[1])} tile words)")
        table = MeasurementTable(self.np_i, self.np_f)
        value = (table, self.planes) if self.return_planes else table
        self._value = (value,)
        self._release()
        return value

    def _release(self):
        if self._slot is not None and self._slot.pending is self:
            self._slot.pending = None
        self._slot = None
        self.hp_i = self.hp_f = self.hp_s = self.np_i = self.np_f = None
        self.planes = None if not self.return_planes else self.planes


def measure_instances(instances: Union[object, Sequence[object]],
                      output_size: Optional[Tuple[int, int]] = None,
                      classes_of_interest: Optional[Sequence[int]] = None, *,
                      mask_threshold: float = 0.5, pixels_per_metric: float = 0.85,
                      image_idx_offset: int = 0, return_planes: bool = False,
                      write_planes: bool = False, gather: bool = False,
                      gather_counts: Optional[Sequence[int]] = None,
                      gather_dst: Optional[int] = None, gather_sink: str = "auto",
                      mask_channel_offset: int = 0,
                      pipeline_chunks: int = 4, device=None, _exact_words: bool = False):
    """Per-instance measurement rows for one image or a batch of images.

    ``instances``: a Detectron2-style ``Instances`` (or a list of them, one per image)
    holding the RAW predictor output -- ``pred_boxes`` in network-input coordinates
    (``image_size``), ``scores``, ``pred_classes`` and ``pred_masks`` as N x 1 x 28 x 28
    probabilities -- on the CPU (pinned memory is used as is) or on the CUDA device.
    ``output_size`` (H, W) is the original image size the boxes are rescaled to
    (``detector_postprocess``); default: each ``image_size``.  All images of one call must
    share the output size.  ``classes_of_interest`` keeps only those classes
    (nn_inference.py:379).  Returns a ``MeasurementTable`` (and the device bit-planes when
    ``return_planes``); an empty selection returns an empty table (the reference prints
    and returns, nn_inference.py:383-385).  ``gather=True`` (under torch.distributed, one
    process per GPU, images sharded over ranks) all-gathers the device rows of every rank
    before the host read, so each rank returns the whole job's table; ``gather_counts``
    (rows per rank, when the caller knows them) skips the count exchange; ``gather_dst=r``
    brings the whole table to the host of rank r only (the other ranks return their own rows).
    With a destination and all ranks on one node the HOST is the sink (``gather_sink="auto"`` /
    ``"host"``): every rank copies its own rows into one host table shared by the ranks
    (``uwcv.dist.SharedHostTable``), nothing is gathered on the devices and no rank reads another
    rank's rows over PCIe; ``gather_sink="device"`` keeps the device all-gather + one big read.
    Tables of host-gathered calls are views of that shared memory: copy or drop them within two
    calls.

    Single-forward form: instead of ``pred_masks`` the instances may carry
    ``pred_mask_logits`` (N x K x 28 x 28, the mask head's raw output); the kernel reads channel
    ``pred_classes[i] + mask_channel_offset`` and applies the sigmoid itself, which is what
    detectron2's ``mask_rcnn_inference`` does before ``detector_postprocess``.

    An optional int field ``orig_idx`` (as ``uwcv.take_instances`` attaches) is reported as the
    ``inst_idx`` column instead of the position inside the call, so that the shards of one
    image (``shard_instances_by_tile``) gather into the table of the whole image.
    """
    return submit_measure_instances(
        instances, output_size, classes_of_interest, mask_threshold=mask_threshold,
        pixels_per_metric=pixels_per_metric, image_idx_offset=image_idx_offset,
        return_planes=return_planes, write_planes=write_planes, gather=gather,
        gather_counts=gather_counts, gather_dst=gather_dst, gather_sink=gather_sink,
        mask_channel_offset=mask_channel_offset, pipeline_chunks=pipeline_chunks, device=device,
        _exact_words=_exact_words).result()


class MeasurementStream:
    """Throughput form of ``measure_instances`` for a sequence of batches (the reference
    loops over the images of a folder, nn_inference.py:485-498): up to ``depth`` calls are
    in flight, so the host->device copy of batch i + 1 and the device->host read of batch
    i - 1 run under the kernels of batch i.

        stream = uwcv.MeasurementStream(device, depth=3)
        for table in stream.map(batches, (H, W)): ...

    ``submit`` returns a ``PendingTable``; tables come back in submission order.

    ``fills``: "per_chunk" starts a call's plane fill as soon as its first chunk of masks has
    landed (shortest path through ONE call: right for depth <= 2, where the next call's copies
    are only issued once this call's predecessor has been collected); "once" writes a call's
    planes with one fill launch behind its last chunk (each launch pays its ramp and drain:
    right for depth >= 3, where the copy engine runs back to back and the previous call's fill
    covers the wait).  Default: by depth.  Planes and rows are the same bytes either way."""

    def __init__(self, device=None, depth: int = 3, fills: Optional[str] = None):
        self.device = _require_cuda(device)
        self.depth = max(1, int(depth))
        if fills not in (None, "per_chunk", "once"):
            raise ValueError("fills must be 'per_chunk' or 'once'")
        self.fill_once = (self.depth >= 3) if fills is None else (fills == "once")
        self._next = 0

    def submit(self, instances, output_size=None, classes_of_interest=None, **kw) -> PendingTable:
        if kw.get("gather") and dist_is_multi():
            from .dist import SharedHostTable
            if self.depth > SharedHostTable.SETS - 1:
                # (a set of the shared host table is written again only when the table that last
                #  used it has been let go of: depth calls in flight + the one the caller reads)
                raise ValueError(f"gathered calls: depth <= {SharedHostTable.SETS - 1}")
        slot = self._next
        self._next = (self._next + 1) % self.depth
        return submit_measure_instances(instances, output_size, classes_of_interest,
                                        device=self.device, _slot=slot,
                                        _fill_once=self.fill_once, **kw)

    def map(self, batches, output_size=None, classes_of_interest=None, **kw):
        inflight: List[PendingTable] = []
        for b in batches:
            inflight.append(self.submit(b, output_size, classes_of_interest, **kw))
            if len(inflight) >= self.depth:
                yield inflight.pop(0).result()
        while inflight:
            yield inflight.pop(0).result()


def _rank_counts(gather_counts, n: int, world: int, dev) -> List[int]:
    """Rows per rank of a gathered call: as given, else exchanged (one synchronising read)."""
    if gather_counts is not None:
        return [int(c) for c in gather_counts]
    import torch.distributed as dist
    c_l = torch.tensor([n], dtype=torch.int64, device=dev)
    c_all = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(c_all, c_l)
    return c_all.cpu().tolist()


def _copy_status(eng, slot, pend, status, d_ok):
    """(On the d2h stream.)  Status word and the validity flag of the device-input path."""
    hp_s = slot.pinned("status", (4,), torch.int64)
    hp_s.copy_(status, non_blocking=True)
    if d_ok is not None:
        pend.hp_ok = slot.pinned("ok", (1,), torch.int64)
        pend.hp_ok.copy_(d_ok, non_blocking=True)
        d_ok.record_stream(eng.d2h_stream)
    return hp_s


def _read_back_to_shared_table(eng, slot, pend, ht, call, dst: int, rows_done, rows_i, rows_f,
                               status, d_ok, n: int) -> "PendingTable":
    """The host is the sink of the gather: this rank's rows go to its offset of the node-shared
    host table; the destination rank returns the whole table, the others their own rows."""
    import torch.distributed as dist
    hk, hseq, hbase, htotal = call
    is_dst = dist.get_rank() == dst
    with torch.cuda.stream(eng.d2h_stream):
        eng.d2h_stream.wait_event(rows_done)
        ht.copy_rows(hk, hseq, hbase, rows_i, rows_f)
        hp_s = _copy_status(eng, slot, pend, status, d_ok)
        done = torch.cuda.Event()
        done.record(eng.d2h_stream)
    lo, hi = (0, htotal) if is_dst else (hbase, hbase + n)
    pend.np_i, pend.np_f = ht.arrays(hk, hseq, lo, hi, whole=is_dst)
    if is_dst:
        pend._after_sync = lambda: ht.wait_all(hk, hseq)
    pend.n = hi - lo
    pend.hp_i = pend.hp_f = None
    pend.hp_s = hp_s
    pend._done = done
    slot.pending = pend
    return pend


def _stage_small_arrays(eng, slot, dev, main, n, boxes, sl, cl, il, jl):
    """The five small per-instance arrays (boxes, scores, classes, image / instance index) in ONE
    packed device buffer of the slot.  Returns their device views."""
    nb = dict(non_blocking=True)
    small_parts = (("boxes", [boxes], torch.float32, 4), ("scores", sl, torch.float32, 1),
                   ("classes", cl, torch.int64, 1), ("img", il, torch.int32, 1),
                   ("inst", jl, torch.int32, 1))
    on_device = any(p_.is_cuda for _, parts, _, _ in small_parts for p_ in parts)
    # one packed buffer for the five small arrays (every segment 256-byte aligned)
    offs, total_b = {}, 0
    for name, _parts, dt, width in small_parts:
        offs[name] = total_b
        total_b += (n * width * dt.itemsize + 255) // 256 * 256
    d_small = slot.device("small", (total_b,), torch.uint8)

    def seg(buf, name, dt, width):
        v = buf[offs[name]: offs[name] + n * width * dt.itemsize].view(dt)
        return v.view(n, width) if width > 1 else v

    d_boxes, d_scores, d_classes, d_img, d_inst = (seg(d_small, nm, dt, w)
                                                   for nm, _p, dt, w in small_parts)
    if on_device:
        # device-resident inputs: plain device copies behind their producer
        eng.small_stream.wait_stream(main)
        with torch.cuda.stream(eng.small_stream):
            for (name, parts, dt, w), dst in zip(small_parts, (d_boxes, d_scores, d_classes,
                                                               d_img, d_inst)):
                src = parts[0] if len(parts) == 1 else torch.cat([p_.to(dev) for p_ in parts])
                dst.copy_(src.to(dt).reshape(dst.shape), **nb)
            ev_small = torch.cuda.Event()
            ev_small.record(eng.small_stream)
        main.wait_event(ev_small)
    else:
        # host inputs: packed into pinned (device-mapped) memory and pulled in by ONE kernel on
        # the main stream (uwcv_ingest).  The copy engine serves copies in issue order: 1 MB of
        # boxes queued behind 200 MB of masks would hold the layout back for milliseconds
        h_small = slot.pinned("small", (total_b,), torch.uint8)
        for name, parts, dt, w in small_parts:
            dst = seg(h_small, name, dt, w)
            if len(parts) > 1:
                torch.cat(parts, out=dst)
            else:
                dst.copy_(parts[0].reshape(dst.shape))
        _lib.check(eng.L.uwcv_ingest(_ptr(h_small), _ptr(d_small), total_b, _stream_ptr(dev)),
                   "uwcv_ingest")
        eng.launches += 1
    return d_boxes, d_scores, d_classes, d_img, d_inst


@dataclass
class _CallInputs:
    """What one call feeds the kernels, still as per-image parts (host or device tensors)."""
    boxes: list            # output-space boxes (scaled, clipped, empty ones dropped)
    scores: list
    classes: list
    masks: list            # [n_k, channels * 784] float32
    image_idx: list
    inst_idx: list
    counts: list           # instances per image
    early: object          # mask copies already under way (_issue_mask_copies), or None
    d_ok: object           # device flag of the optimistic device-input path, or None
    logits: bool
    channels: int


def _collect_call_inputs(eng, slot, dev, batch, H, W, output_size, classes_of_interest,
                         image_idx_offset, pipeline_chunks, no_fast) -> _CallInputs:
    """detector_postprocess's box handling (scale, clip, drop empty) for every image of the call
    and the per-instance index columns.  Three paths with the same result: host tensors with one
    shared image size (one vectorised pass, the mask copies start first), device-resident
    predictor output (no host synchronisation: validity is a device flag), and the general
    per-image path (class filter, mixed sizes, ``orig_idx``)."""
    bl, sl, cl, ml, il, jl = [], [], [], [], [], []
    mfields = [_mask_field(inst) for inst in batch]
    logits = mfields[0][1]
    channels = int(mfields[0][0].shape[1]) if mfields[0][0].dim() == 4 else 1
    for m, lg in mfields:
        if lg != logits or (len(m) and int(m.shape[1]) != channels):
            raise ValueError("all images of one call must carry the same kind of mask field")
        if len(m) and tuple(m.shape[-2:]) != (MASK_SIDE, MASK_SIDE):
            raise ValueError(f"uwcv is built for {MASK_SIDE}x{MASK_SIDE} mask heads, got {tuple(m.shape)}")
    mrow = channels * MASK_SIDE * MASK_SIDE
    sizes = {tuple(int(v) for v in inst.image_size) for inst in batch}
    if output_size is None and len(sizes) > 1:
        raise ValueError("all images of one call must share the output size")
    early = None
    counts_fast = None
    if len(sizes) == 1 and classes_of_interest is None and len(batch) > 1 and \
            all(not _as_box_tensor(i.pred_boxes).is_cuda for i in batch):
        # fast host path: one scale / clip / non-empty over the concatenated boxes; the mask
        # probabilities start moving to the device first (optimistically: no box is dropped)
        lens = [len(i) for i in batch]
        ml0 = [m.to(torch.float32).reshape(-1, mrow) for m, _ in mfields]
        early = _issue_mask_copies(eng, slot, dev, ml0, lens, pipeline_chunks)
        allb = torch.cat([_as_box_tensor(i.pred_boxes) for i in batch])
        b_all, keep_all = scale_clip_boxes(allb, batch[0].image_size, (H, W))
        if bool(keep_all.all()):
            bl = [b_all]
            sl = [i.scores.to(torch.float32) for i in batch]
            cl = [i.pred_classes.to(torch.int64) for i in batch]
            ml = ml0
            lt = torch.tensor(lens, dtype=torch.int64)
            il = [torch.repeat_interleave(
                torch.arange(len(batch), dtype=torch.int32) + image_idx_offset, lt)]
            offs = torch.cumsum(lt, 0) - lt
            if all(_has(i, "orig_idx") for i in batch):          # shards keep their global numbering
                jl = [torch.cat([i.orig_idx.to(torch.int32).cpu() for i in batch])]
            else:
                jl = [(torch.arange(int(lt.sum()), dtype=torch.int64)
                       - torch.repeat_interleave(offs, lt)).to(torch.int32)]
            counts_fast = lens
        else:
            early = None
    d_ok = None
    if counts_fast is None and not no_fast and len(sizes) == 1 and classes_of_interest is None and \
            all(_as_box_tensor(i.pred_boxes).is_cuda and _as_box_tensor(i.pred_boxes).device == dev
                for i in batch) and not any(_has(i, "orig_idx") for i in batch):
        # device-resident predictor output (the single-forward path): nothing here may wait for
        # the GPU, so "no box is dropped, every box is finite" is ASSUMED, computed on the device
        # as a flag and read back with the rows; a call whose flag is false is repeated through
        # the general path below (rare: empty or non-finite boxes after clipping)
        lens = [len(i) for i in batch]
        allb = torch.cat([_as_box_tensor(i.pred_boxes) for i in batch]).to(torch.float32)
        b_all, keep_all = scale_clip_boxes(allb, batch[0].image_size, (H, W), check=False)
        d_ok = (keep_all.all() & torch.isfinite(allb).all()).to(torch.int64).reshape(1)
        bl = [b_all]
        sl = [i.scores.to(torch.float32) for i in batch]
        cl = [i.pred_classes.to(torch.int64) for i in batch]
        ml = [m.to(torch.float32).reshape(-1, mrow) for m, _ in mfields]
        lt = torch.tensor(lens, dtype=torch.int64)
        il = [torch.repeat_interleave(torch.arange(len(batch), dtype=torch.int32) + image_idx_offset, lt)]
        offs = torch.cumsum(lt, 0) - lt
        jl = [(torch.arange(int(lt.sum()), dtype=torch.int64) - torch.repeat_interleave(offs, lt)).to(torch.int32)]
        counts_fast = lens
    for k, inst in enumerate(batch if counts_fast is None else []):
        boxes, scores, classes, masks = _gather_fields(inst, classes_of_interest)
        out_sz = (H, W)
        if output_size is None and tuple(int(v) for v in inst.image_size) != out_sz:
            raise ValueError("all images of one call must share the output size")
        b, keep = scale_clip_boxes(boxes, inst.image_size, out_sz)
        nk = int(keep.sum().item()) if keep.numel() else 0
        if nk != keep.numel():
            b, scores, classes, masks = b[keep], scores[keep], classes[keep], masks[keep]
        bl.append(b)
        sl.append(scores.to(torch.float32))
        cl.append(classes.to(torch.int64))
        ml.append(masks.to(torch.float32).reshape(-1, mrow))
        il.append(torch.full((nk,), image_idx_offset + k, dtype=torch.int32))
        if _has(inst, "orig_idx"):
            oi = inst.orig_idx.to(torch.int32).cpu()
            if classes_of_interest is not None:
                sel = torch.zeros(len(oi), dtype=torch.bool)
                for c in classes_of_interest:
                    sel |= inst.pred_classes.cpu() == int(c)
                oi = oi[sel]
            jl.append(oi[keep.cpu()] if nk != keep.numel() else oi)
        else:
            jl.append(torch.arange(nk, dtype=torch.int32))
    counts = counts_fast if counts_fast is not None else [int(b.shape[0]) for b in bl]
    return _CallInputs(bl, sl, cl, ml, il, jl, counts, early, d_ok, logits, channels)


def submit_measure_instances(instances, output_size=None, classes_of_interest=None, *,
                             mask_threshold: float = 0.5, pixels_per_metric: float = 0.85,
                             image_idx_offset: int = 0, return_planes: bool = False,
                             write_planes: bool = False, gather: bool = False,
                             gather_counts: Optional[Sequence[int]] = None,
                             gather_dst: Optional[int] = None, gather_sink: str = "auto",
                             mask_channel_offset: int = 0,
                             pipeline_chunks: int = 4, device=None, _exact_words: bool = False,
                             _slot: int = 0, _no_fast: bool = False,
                             _fill_once: bool = False) -> PendingTable:
    """Enqueue one ``measure_instances`` call (same arguments) and return its handle."""
    single = not isinstance(instances, (list, tuple))
    batch: List[object] = [instances] if single else list(instances)
    dev = _require_cuda(device)
    eng = Engine.get(dev)
    empty = (MeasurementTable.empty(), None) if return_planes else MeasurementTable.empty()
    if not batch:
        return PendingTable.ready(MeasurementTable.empty())
    H, W = (int(output_size[0]), int(output_size[1])) if output_size is not None else \
        tuple(int(v) for v in batch[0].image_size)
    slot = eng.slot(_slot)
    if slot.pending is not None:            # the slot's previous call has not been collected
        slot.pending.result()

    inp = _collect_call_inputs(eng, slot, dev, batch, H, W, output_size, classes_of_interest,
                               image_idx_offset, pipeline_chunks, _no_fast)
    bl, sl, cl, ml, il, jl = inp.boxes, inp.scores, inp.classes, inp.masks, inp.image_idx, inp.inst_idx
    early, d_ok, logits, channels = inp.early, inp.d_ok, inp.logits, inp.channels
    boxes = torch.cat(bl)
    n = int(boxes.shape[0])
    gathered = gather and dist_is_multi()
    if n == 0 and not gathered:
        return PendingTable.ready(empty)
    # Workspace sizing: the exact tile-word count is only computed for the first call (or
    # after an overflow); afterwards the cached capacity is reused and the device status
    # word reports an overflow, in which case the call is repeated with the exact size.
    # (a gathered call must not be repeated on one rank only: it is sized by an upper bound)
    if _exact_words:
        n_words = tile_words(boxes, H, W)
    elif eng._ws is None or (gathered and not boxes.is_cuda):
        n_words = tile_words_bound(boxes)             # (device boxes: one synchronising read, first call only)
    else:
        n_words = eng._cap_words
    counts = inp.counts
    main = torch.cuda.current_stream(dev)
    nb = dict(non_blocking=True)

    def retry():
        return measure_instances(
            instances, output_size, classes_of_interest, mask_threshold=mask_threshold,
            pixels_per_metric=pixels_per_metric, image_idx_offset=image_idx_offset,
            return_planes=return_planes, write_planes=write_planes, gather=gather,
            gather_counts=gather_counts, gather_dst=gather_dst, gather_sink=gather_sink,
            mask_channel_offset=mask_channel_offset, pipeline_chunks=pipeline_chunks,
            device=device, _exact_words=True)

    pend = PendingTable(slot, None if _exact_words else retry)
    pend.return_planes = return_planes
    if d_ok is not None:
        def redo():
            if gathered:
                raise RuntimeError("uwcv: empty or non-finite boxes in a gathered call with device-resident "
                                   "inputs; filter them (Boxes.nonempty) before the call")
            return submit_measure_instances(
                instances, output_size, classes_of_interest, mask_threshold=mask_threshold,
                pixels_per_metric=pixels_per_metric, image_idx_offset=image_idx_offset,
                return_planes=return_planes, write_planes=write_planes,
                mask_channel_offset=mask_channel_offset, pipeline_chunks=pipeline_chunks,
                device=device, _no_fast=True).result()
        pend._redo = redo
    with torch.cuda.device(dev):
        if early is None:
            early = _issue_mask_copies(eng, slot, dev, ml, counts, pipeline_chunks)
        d_masks, ev_in, bounds, resume_mask_copies = early
        rows_i = slot.device("rows_i", (n, NUM_INT), torch.int64)
        rows_f = slot.device("rows_f", (n, NUM_FLOAT), torch.float64)
        status = slot.device("status", (4,), torch.int64)
        if return_planes:
            planes = eng.alloc_planes(n, H, W)            # handed to the caller
        elif write_planes:
            planes = eng.scratch_planes(n, H, W)          # engine-owned, reused across calls
        else:
            planes = None
        pend.planes = planes if return_planes else None
        d_boxes, d_scores, d_classes, d_img, d_inst = _stage_small_arrays(
            eng, slot, dev, main, n, boxes, sl, cl, il, jl)
        resume_mask_copies()
        ws = eng._workspace(n, n_words)
        common = dict(image_idx=d_img, inst_idx=d_inst, classes=d_classes, scores=d_scores,
                      threshold=mask_threshold, pixels_per_metric=pixels_per_metric,
                      planes=planes, n_tile_words=n_words, rows_i=rows_i, rows_f=rows_f,
                      status=status, mask_channels=channels,
                      channel_offset=mask_channel_offset if channels > 1 else 0, logits=logits)
        # layout for the whole call as soon as the boxes are on the device; paste chunk by chunk
        # as the mask probabilities arrive; one border-trace launch over all instances (its
        # duration is set by the longest serial chain, not by the instance count)
        r = n
        out = {}

        def gather_rows():
            from .dist import all_gather_table
            out["i"], out["f"] = all_gather_table(rows_i, rows_f, counts=gather_counts)

        if gather_sink not in ("auto", "host", "device"):
            raise ValueError("gather_sink must be 'auto', 'host' or 'device'")
        ht = None
        if gathered and gather_dst is not None and gather_sink != "device":
            ht = eng.host_table()
            if ht is None and gather_sink == "host":
                raise RuntimeError("uwcv: the shared host table is not available for this group")
        if ht is not None:
            from .dist import SharedTableUnavailable
            try:
                hk, hseq, hbase, htotal = ht.begin(_rank_counts(gather_counts, n, ht.world, dev))
            except SharedTableUnavailable as e:       # (raised by all ranks together)
                if gather_sink == "host":
                    raise
                import warnings
                warnings.warn(f"uwcv: {e}; gathering on the devices")
                eng._host_table = False
                ht = None
        fg = eng.fused_gather() if (gathered and ht is None) else None
        gstruct = gset = None
        if fg is not None:
            # fused all-gather: the trace kernel stores the rows into every rank's table
            # (symmetric memory over NVLink), a signal barrier completes them -- no collective
            # kernel has to find SMs next to the following call's paste
            gstruct, gset, g_total = fg.begin(_rank_counts(gather_counts, n, fg.world, dev), dst=gather_dst)
        if n > 0:
            rows_done = eng.run_overlapped(
                d_masks, d_boxes, H, W,
                paste_ranges=[(lo, hi - lo, ev_in[c]) for c, (i0, i1, lo, hi) in enumerate(bounds)],
                gather=gstruct, after=(lambda: fg.barrier(gset)) if fg is not None else None,
                fill_once=_fill_once, **common)
            if planes is not None and eng.planes_done is not None:
                pend._planes_done = eng.planes_done
        else:
            status.zero_()
            with torch.cuda.stream(eng.trace_stream):
                eng.trace_stream.wait_stream(main)
                if fg is not None:
                    fg.barrier(gset)
                rows_done = torch.cuda.Event()
                rows_done.record(eng.trace_stream)
        if fg is not None:
            out["i"], out["f"] = fg.tables(gset, g_total)
        elif ht is not None:
            pass                                   # the host is the sink: nothing moves between devices
        elif gathered:
            # NCCL fallback: the collective stays on the main stream, behind the trace: an NCCL
            # kernel issued from the trace stream would have to wait for SMs held by the next
            # call's paste (measured: 8.2 vs 7.5 ms/step at 2 GPUs), so gathered calls run back
            # to back
            main.wait_event(rows_done)
            gather_rows()
            rows_done = torch.cuda.Event()
            rows_done.record(main)
        out_i, out_f = (out["i"], out["f"]) if (gathered and ht is None) else (rows_i, rows_f)
        if ht is not None:
            return _read_back_to_shared_table(eng, slot, pend, ht, (hk, hseq, hbase, htotal), int(gather_dst),
                                              rows_done, rows_i, rows_f, status, d_ok, n)
        if gathered and gather_dst is not None:
            import torch.distributed as dist
            if dist.get_rank() != int(gather_dst):
                out_i, out_f = rows_i, rows_f          # only the destination reads the whole table
        r = int(out_i.shape[0])
        # the rows land in pinned memory that the returned table owns (no host copy); the
        # engine recycles the buffer when the table is gone
        hp_i, hp_f, pend.np_i, pend.np_f = eng.host_rows(r)
        with torch.cuda.stream(eng.d2h_stream):
            eng.d2h_stream.wait_event(rows_done)
            hp_i.copy_(out_i, **nb)
            hp_f.copy_(out_f, **nb)
            hp_s = _copy_status(eng, slot, pend, status, d_ok)
            done = torch.cuda.Event()
            done.record(eng.d2h_stream)
        if fg is not None:
            fg.read_done[gset] = done
        elif gathered and out_i is not rows_i:   # gathered tensors: the main stream's pool
            out_i.record_stream(eng.d2h_stream)
            out_f.record_stream(eng.d2h_stream)
        # rows / status / inputs are per slot and a slot is only reused after its call has
        # been collected, so the main stream never waits for the device->host read
    pend.n = r
    pend.hp_i, pend.hp_f, pend.hp_s = hp_i, hp_f, hp_s
    pend._done = done
    slot.pending = pend
    return pend


def _issue_mask_copies(eng: "Engine", slot: _Slot, dev, ml, counts, pipeline_chunks: int):
    """Enqueue the host->device copy of the FIRST chunks of mask probabilities on the engine's
    copy stream and return (device masks, one event per chunk, chunk bounds, resume); calling
    ``resume()`` enqueues the remaining chunks (none today: the small arrays -- boxes, scores,
    ... -- are pulled in by a kernel, so all mask chunks can be queued at once; the copy engine
    serves copies in issue order, and boxes queued behind 200 MB of masks would hold the layout,
    and with it every paste, back until the last mask has landed)."""
    n = int(sum(counts))
    nchunks = max(1, min(int(pipeline_chunks), len(ml)))
    if all(m.is_cuda for m in ml):
        nchunks = 1          # nothing crosses PCIe: chunks would only split the kernels (4.47 vs 5.02 ms per
        #                      64 000 instances, profiles/r02_e2e_depth.txt)
    per = (len(ml) + nchunks - 1) // nchunks
    bounds = []                                   # (first image, last image + 1, lo row, hi row)
    lo = 0
    for c in range(nchunks):
        i0, i1 = c * per, min(len(ml), (c + 1) * per)
        if i0 >= i1:
            break
        hi = lo + sum(counts[i0:i1])
        bounds.append((i0, i1, lo, hi))
        lo = hi
    starts = [0]
    for k in counts:
        starts.append(starts[-1] + k)
    main = torch.cuda.current_stream(dev)
    mrow = int(ml[0].shape[1]) if len(ml) else MASK_SIDE * MASK_SIDE
    # device-resident masks that already sit back to back in one allocation (the mask head's
    # output split per image) are used where they are
    live = [m for m in ml if m.shape[0] > 0]
    if live and all(m.is_cuda and m.device == dev and m.is_contiguous() and
                    m.dtype == torch.float32 for m in live) and \
            all(a.data_ptr() + a.numel() * 4 == b.data_ptr() and
                a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
                for a, b in zip(live[:-1], live[1:])) and live[0].data_ptr() % 16 == 0:
        whole = torch.as_strided(live[0], (n, mrow), (mrow, 1))
        ev = torch.cuda.Event()
        ev.record(main)
        return whole, [ev for _ in bounds], bounds, (lambda: None)
    with torch.cuda.device(dev):
        d_masks = slot.device("masks", (n, mrow), torch.float32)
        ev_in = [torch.cuda.Event() for _ in bounds]
        # the slot's previous call (the only other user of this buffer) was collected before
        # this one was admitted, so the copies need not wait for the main stream
        if any(m.is_cuda for m in ml):              # device-resident inputs: after their producer
            eng.h2d_stream.wait_stream(main)
            for m in ml:
                if m.is_cuda:
                    m.record_stream(eng.h2d_stream)

        def issue(c0: int, c1: int) -> None:
            with torch.cuda.device(dev), torch.cuda.stream(eng.h2d_stream):
                for c in range(c0, c1):
                    i0, i1, _lo, _hi = bounds[c]
                    for i in range(i0, i1):         # pinned sources stay pinned: async copies
                        if counts[i]:
                            d_masks[starts[i]:starts[i + 1]].copy_(ml[i], non_blocking=True)
                    ev_in[c].record(eng.h2d_stream)

        # all chunks go out at once: the small arrays do not travel through the copy engine
        # (uwcv_ingest), so nothing the first kernel needs queues behind them
        n_early = len(bounds)
        issue(0, n_early)
    return d_masks, ev_in, bounds, (lambda: issue(n_early, len(bounds)))


def dist_is_multi() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def get_counts(instances) -> List[int]:
    """What GetCounts (nn_inference.py:355-366) is meant to produce: instances per class id
    0..3 (the reference compares against ids 1..4 and duplicates ``classes == 3``)."""
    from .schema import CLASS_NAMES
    c = instances.pred_classes
    c = c.cpu().numpy() if hasattr(c, "cpu") else np.asarray(c)
    return [int((c == k).sum()) for k in range(len(CLASS_NAMES))]
