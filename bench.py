#!/usr/bin/env python
"""Benchmark of the uw-com-vision post-inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path

Workload at N = 1 (BASELINE.json configs[1]): a batch of 64 synthetic 2048 x 2048 images
with 1000 Mask R-CNN-shaped instances each (28 x 28 mask probabilities, boxes, scores,
classes; uwcv/synth.py).  At N > 1 (configs[2], as written): 256 such images, image b on rank
b % N (strong scaling: 128 / 64 / 32 images per GPU at 2 / 4 / 8), with the weak-scaling
measurement (64 images per GPU) reported beside it as ``weak_scaling``; ``--images`` forces a
fixed number of images per GPU.  One step = one pass of the hot path over the rank's batch:
tile layout -> paste / threshold / bit-pack / moments into tiles -> [Detectron2-literal full-frame
planes (1 bit / pixel) written from the tiles || border trace + descriptors]; at N > 1 ranks the
gather of the measurement table follows.

Printed JSON keys (driver contract): metric/value/unit (instances measured per second,
whole job), ms_per_step, mp_per_sec, e2e (same metric through the public call,
uwcv.MeasurementStream.map = measure_instances with two calls in flight, with pinned HOST
inputs: H2D of every step's inputs and D2H of its rows inside the timed region; the
synchronous one-call-per-step time is reported beside it), roofline (plane-fill kernel, plane bytes / its live
CUDA-event time inside the timed region vs measured HBM peak), cpu_baseline (the reference's CPU path on a bounded sample, N = 1 only),
gpu_launches, clocks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "uw-com-vision_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H = W = 2048
IMAGES_PER_GPU = 64
CONFIG2_IMAGES = 256
INSTANCES_PER_IMAGE = 1000
METRIC = "instances_measured_per_sec"
UNIT = "instances/s"
WORKLOAD = "configs[1]: 64 synthetic 2048x2048 images x 1000 instances per GPU, full-frame bit-planes"
WORKLOAD2 = ("configs[2]: 256 synthetic 2048x2048 images x 1000 instances, image b on rank b % N, "
             "full-frame bit-planes, gather of the measurement table")


def algorithmic_bytes_per_instance(h: int, w: int) -> int:
    """SURVEY.md 8(d), full-frame contract: read 3136 (probs) + 16 (box) + 4 (score) +
    8 (class) + 4 (image idx); write H*W/8 (bit-plane) + 400 (row)."""
    return 3168 + h * w // 8 + 400


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def window(self, t0: float, t1: float):
        """Samples whose nvidia-smi timestamp falls inside [t0, t1] (epoch seconds)."""
        import datetime
        out = []
        for ln in list(self.lines):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                continue
            if t0 - 0.03 <= ts <= t1 + 0.03:
                out.append(f)
        return out

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def summary(self, fields, window: str):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        for f in fields:
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path
# ---------------------------------------------------------------------------------------

def reference_step(inst, out_size, literal_paint=True):
    """What the reference script does per image after the forward pass: Detectron2
    post-process (paste all N masks to N x H x W bool, nn_inference.py:372) once, then
    GetMask_Contours for each of the four class keywords (nn_inference.py:487-496) with the
    literal union-paint loop (:399-401).  (The script repeats the post-process 12 times per
    image; one is charged here.)  Returns the number of instances processed."""
    from oracle import d2, measure as M, pipeline as P
    o = P.to_oracle_instances(inst)
    res = d2.detector_postprocess(o, out_size[0], out_size[1], 0.5)
    classes = res.pred_classes.numpy()
    masks = res.pred_masks.numpy()
    rows = 0
    for k in range(len(M.KEYWORDS)):
        try:
            r = M.get_mask_contours((out_size[0], out_size[1], 3), classes, masks, [k],
                                    literal_paint=literal_paint)
        except ValueError:
            r = None
        rows += 0 if r is None else len(r)
    return len(res), rows


def cpu_sample_size(steps: int, warmup: int) -> int:
    return int(min(INSTANCES_PER_IMAGE, max(25, INSTANCES_PER_IMAGE * 3 // max(steps + warmup, 1))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from uwcv import synth
    n = cpu_sample_size(args.steps, args.warmup)
    inst = synth.blob_instances(0, n, H, W, seed=1234)
    for _ in range(args.warmup):
        reference_step(inst, (H, W))
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        k, _rows = reference_step(inst, (H, W))
        done += k
    dt = time.perf_counter() - t0
    val = done / dt
    sample = f"1 image 2048x2048 x {n} instances per step (detector_postprocess + 4 x GetMask_Contours)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "mp_per_sec": args.steps * H * W / 1e6 / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------

class DeviceBatch:
    """Device-resident inputs + outputs of one rank's batch (the `value` leg)."""

    def __init__(self, eng, batch, rank, world, dev, with_planes=True):
        import uwcv
        from uwcv import api
        self.eng, self.dev = eng, dev
        boxes = torch.cat([api.scale_clip_boxes(b.pred_boxes.tensor, (H, W), (H, W))[0] for b in batch])
        self.n = n = int(boxes.shape[0])
        self.words = api.tile_words(boxes, H, W)
        self.boxes = boxes.to(dev)
        self.masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous().to(dev)
        self.scores = torch.cat([b.scores for b in batch]).to(dev)
        self.classes = torch.cat([b.pred_classes for b in batch]).to(dev)
        self.img = torch.cat([torch.full((len(b),), rank + i * world, dtype=torch.int32)
                              for i, b in enumerate(batch)]).to(dev)
        self.inst = torch.cat([torch.arange(len(b), dtype=torch.int32) for b in batch]).to(dev)
        self.planes = eng.alloc_planes(n, H, W) if with_planes else None
        # two row tables / status words, used in turn: the border trace of step i runs on the
        # engine's trace stream under the paste of step i + 1 (Engine.run_overlapped)
        self.rows = [(torch.empty((n, 20), dtype=torch.int64, device=dev),
                      torch.empty((n, 30), dtype=torch.float64, device=dev)) for _ in range(2)]
        self.stat = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(2)]
        self.tick = 0

    def args(self, planes=True):
        return dict(image_idx=self.img, inst_idx=self.inst, classes=self.classes, scores=self.scores,
                    planes=self.planes if planes else None, n_tile_words=self.words)

    def stage(self, stages):                 # one kernel group alone, for the per-kernel times
        self.eng.run(self.masks, self.boxes, H, W, rows_i=self.rows[0][0], rows_f=self.rows[0][1],
                     stages=stages, **self.args())

    def step(self, gather=None, after=None, planes=True):
        k = self.tick & 1
        self.tick += 1
        ri, rf = self.rows[k]
        self.eng.run_overlapped(self.masks, self.boxes, H, W, rows_i=ri, rows_f=rf, status=self.stat[k],
                                gather=gather, after=after, **self.args(planes))
        return k

    def check_status(self):
        for st in self.stat:
            if int(st.cpu()[0]) != 0:
                raise RuntimeError(f"workspace overflow: {st.cpu().tolist()}")


def timed_steps(db, steps, warmup, barrier, step_fn):
    """W warm-up steps, then exactly K steps between CUDA events on the launching stream, a
    barrier + synchronize on both sides.  Returns (ms, wall0, wall1)."""
    main = torch.cuda.current_stream(db.dev)
    for _ in range(warmup):
        step_fn()
    barrier()
    db.check_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(steps):
        step_fn()
    main.wait_stream(db.eng.trace_stream)       # the last traces (and gathers) end the region
    main.wait_stream(db.eng.fill_stream)        # ... and the last plane fill (split pipeline)
    e1.record()
    barrier()
    wall1 = time.time()
    return e0.elapsed_time(e1), wall0, wall1


def library_paste_cuda(masks, boxes, img_h, img_w, threshold=0.5):
    """The library composition the reference triggers on a GPU, written with the
    torch calls Detectron2's CUDA branch makes (layers/mask_ops.py paste_masks_in_image /
    _do_paste_mask with skip_empty=False): chunks of at most 1 GiB of float32 result, a
    materialised (n, H, W, 2) grid, F.grid_sample, `>=`, an indexed copy into N x H x W bool.
    Here as a MEASURED BASELINE of the same B200, never as a product path."""
    import torch.nn.functional as F
    n = int(masks.shape[0])
    dev = masks.device
    chunks = torch.chunk(torch.arange(n, device=dev), int(np.ceil(n * img_h * img_w * 4 / 1024 ** 3)))
    img_masks = torch.zeros(n, img_h, img_w, device=dev, dtype=torch.bool)
    for inds in chunks:
        m, b = masks[inds, None, :, :], boxes[inds]
        x0, y0, x1, y1 = torch.split(b, 1, dim=1)
        img_y = torch.arange(0, img_h, device=dev, dtype=torch.float32) + 0.5
        img_x = torch.arange(0, img_w, device=dev, dtype=torch.float32) + 0.5
        img_y = (img_y - y0) / (y1 - y0) * 2 - 1
        img_x = (img_x - x0) / (x1 - x0) * 2 - 1
        k = int(m.shape[0])
        gx = img_x[:, None, :].expand(k, img_y.size(1), img_x.size(1))
        gy = img_y[:, :, None].expand(k, img_y.size(1), img_x.size(1))
        grid = torch.stack([gx, gy], dim=3)
        chunk = F.grid_sample(m, grid.to(m.dtype), align_corners=False)[:, 0]
        img_masks[(inds,)] = (chunk >= threshold).to(dtype=torch.bool)
    return img_masks


def library_baseline_gpu(eng, dev, n_inst=1000, reps=3):
    """One image of the bench workload through the libraries on this GPU: Detectron2's CUDA paste
    (library_paste_cuda), `.to("cpu")` of the N x H x W bool as nn_inference.py:376 reads it, and
    torchvision's CUDA batched_nms on the configs[3] candidates next to uwcv_nms.  Device times from
    CUDA events, the copy by wall clock.  The pixels are compared with this repo's planes for the same
    instances and the count of differing pixels is reported (CUDA grid_sample is not the parity oracle:
    the bar is the reference's CPU path, tests/)."""
    import torchvision
    from uwcv import api, synth
    out = {}
    inst = synth.blob_instances(0, n_inst, H, W, seed=1234)
    bx, keep = api.scale_clip_boxes(inst.pred_boxes.tensor, (H, W), (H, W))
    masks = inst.pred_masks[keep, 0].contiguous().to(dev)
    boxes = bx[keep].contiguous().to(dev)
    n = int(boxes.shape[0])
    bits = library_paste_cuda(masks, boxes, H, W)              # warm-up (allocator, cuDNN-free path)
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(reps):
        del bits
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); bits = library_paste_cuda(masks, boxes, H, W); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    paste_ms = statistics.median(ts)
    t0 = time.perf_counter()
    host = bits.to("cpu").numpy()
    d2h_ms = (time.perf_counter() - t0) * 1e3
    del host
    # this repo's path on the same instances: planes + rows, device-resident inputs
    words = api.tile_words(boxes.cpu(), H, W)
    planes = eng.alloc_planes(n, H, W)
    ri = torch.empty((n, 20), dtype=torch.int64, device=dev)
    rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
    kw = dict(planes=planes, n_tile_words=words, rows_i=ri, rows_f=rf)
    for _ in range(3):
        eng.run(masks, boxes, H, W, **kw)
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.run(masks, boxes, H, W, **kw); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ours_ms = statistics.median(ts)
    mine = eng.unpack(planes, H, W)
    differing = int((mine != bits).sum().item())
    area_equal = bool(torch.equal(bits.flatten(1).sum(1), ri[:, 5]))
    del mine, bits, planes
    out["paste"] = {
        "sample": f"1 image {H}x{W} x {n} instances, device-resident inputs",
        "library_ms": paste_ms, "library_instances_per_s": n / paste_ms * 1e3,
        "library": f"torch {torch.__version__} F.grid_sample + >= + indexed copy, 1 GiB chunks "
                   "(Detectron2 paste_masks_in_image, CUDA branch) -> N x H x W bool",
        "library_mask_d2h_ms": d2h_ms,
        "library_mask_d2h_note": "N x H x W bool .to('cpu') (pageable), what nn_inference.py:376 does next; "
                                 "the contour / descriptor work of the reference then runs on the host (cpu_baseline)",
        "ours_ms": ours_ms, "ours_instances_per_s": n / ours_ms * 1e3,
        "ours": "single-stream uwcv_paste_measure call: planes (bit-packed) + moments + contours + "
                "descriptor rows for the same instances (more work than the library leg: it also measures)",
        "speedup_device": paste_ms / ours_ms,
        "pixels_differing": differing, "pixels_total": n * H * W,
        "area_px_equal_to_library_popcount": area_equal,
    }
    # NMS: fast_rcnn_inference_single_image's filter + batched_nms + top-k on the configs[3] candidates
    cb, cs, cc = synth.clustered_candidates(5000, 4096, 4096, seed=99)
    dcb, dcs, dcc = cb.to(dev), cs.to(dev), cc.to(dev)

    def tv():
        f = dcs > 0.05
        b2, s2, c2 = dcb[f], dcs[f], dcc[f]
        k = torchvision.ops.batched_nms(b2, s2, c2, 0.5)[:6000]
        return f.nonzero()[:, 0][k]

    for _ in range(3):
        kept = tv()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kept = tv(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    tv_ms = statistics.median(ts)
    for _ in range(3):
        keep_o, cnt = eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); keep_o, cnt = eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    out["nms"] = {
        "sample": f"configs[3]: {len(cb)} candidates, 4 classes, score > 0.05, IoU 0.5, top 6000",
        "library_ms": tv_ms, "library": f"torchvision {torchvision.__version__} batched_nms (CUDA) after the score filter",
        "ours_ms": statistics.median(ts), "speedup_device": tv_ms / statistics.median(ts),
        "keep_list_equal": bool(torch.equal(keep_o[: int(cnt[0])].to(torch.int64), kept.to(torch.int64))),
    }
    return out


def other_configs(eng, dev):
    """configs[0] and configs[3] (parity-test shapes, not bench lines) with their kernel breakdown:
    device-resident, full-frame planes, CUDA events per kernel group."""
    import uwcv
    from uwcv import api, synth

    def breakdown(masks, boxes, Hc, Wc, reps=10):
        n = int(boxes.shape[0])
        words = api.tile_words(boxes.cpu(), Hc, Wc)
        planes = eng.alloc_planes(n, Hc, Wc)
        ri = torch.empty((n, 20), dtype=torch.int64, device=dev)
        rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
        kw = dict(planes=planes, n_tile_words=words, rows_i=ri, rows_f=rf)
        for _ in range(3):
            eng.run(masks, boxes, Hc, Wc, **kw)
        t = {1: [], 2: [], 4: []}
        for _ in range(reps):
            for st in (1, 2, 4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); eng.run(masks, boxes, Hc, Wc, stages=st, **kw); b.record(); b.synchronize()
                t[st].append(a.elapsed_time(b))
        lay, pas, tra = (statistics.mean(t[k]) for k in (1, 2, 4))
        tot = lay + pas + tra
        bpi = algorithmic_bytes_per_instance(Hc, Wc)
        del planes
        return {"instances": n, "image": f"{Hc}x{Wc}",
                "kernel_ms": {"layout": lay, "paste_measure": pas, "contour": tra}, "ms": tot,
                "instances_per_s": n / tot * 1e3, "paste_gbs_algorithmic": n * bpi / (pas * 1e-3) / 1e9,
                "bytes_per_instance": bpi}

    out = {}
    gp = os.path.join(ROOT, "tests", "golden", "c1_maskrcnn.npz")
    if os.path.exists(gp):
        g = np.load(gp)
        bx, keep = api.scale_clip_boxes(torch.from_numpy(g["boxes"]), (1024, 1024), (1024, 1024))
        m = torch.from_numpy(g["masks"])[keep, 0].contiguous()
        out["configs[0]_raw_maskrcnn_heads"] = breakdown(m.to(dev), bx[keep].contiguous().to(dev), 1024, 1024)
        out["configs[0]_raw_maskrcnn_heads"]["note"] = (
            "200 raw random-init head outputs: sub-pixel and 1024-px boxes, up to 90 speckle contours "
            "per mask -- an edge-case parity input, one CTA / one trace lane per instance")
    inst = synth.blob_instances(0, 200, 1024, 1024, seed=1234)
    bx, keep = api.scale_clip_boxes(inst.pred_boxes.tensor, (1024, 1024), (1024, 1024))
    out["configs[0]_shape_blob_source"] = breakdown(inst.pred_masks[keep, 0].contiguous().to(dev),
                                                    bx[keep].contiguous().to(dev), 1024, 1024)
    # configs[3]: 4096 x 4096, ~20 k clustered candidates -> score filter + per-class NMS -> ~5 k instances
    Hc = Wc = 4096
    cb, cs, cc = synth.clustered_candidates(5000, Hc, Wc, seed=99)
    dcb, dcs, dcc = cb.to(dev), cs.to(dev), cc.to(dev)
    for _ in range(3):
        keep, cnt = eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000)
    tn = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); keep, cnt = eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000); b.record(); b.synchronize()
        tn.append(a.elapsed_time(b))
    k = int(cnt[0])
    keep = keep[:k].cpu()
    gg = torch.Generator().manual_seed(8)
    rec = breakdown(synth.blob_probs(k, gg).to(dev), cb[keep].contiguous().to(dev), Hc, Wc, reps=5)
    rec["candidates"] = int(len(cb))
    rec["kernel_ms"]["nms"] = statistics.mean(tn)
    out["configs[3]_4096_dense"] = rec
    return out


def run_ours(args):
    import torch.distributed as dist
    import uwcv
    from uwcv import api, dist as udist, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        # the all-gather of step i is issued behind the trace of step i while the paste of step
        # i + 1 already holds the SMs: its kernel must be placed first when CTAs retire
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)

    # N = 1: configs[1].  N > 1: configs[2] as written (256 images over the ranks, strong scaling)
    # unless --images fixes the images per GPU (weak scaling)
    strong = world > 1 and args.images is None and CONFIG2_IMAGES % world == 0
    n_img = CONFIG2_IMAGES // world if strong else (args.images or IMAGES_PER_GPU)
    n_inst = args.instances
    # rank r owns images r, r + world, ... of the (n_img * world)-image set
    batch = synth.blob_batch(n_img, n_inst, H, W, seed=1234, first_image=rank, stride=world)
    for inst in batch:                                  # pinned host copies for the e2e leg
        for k, v in list(inst.get_fields().items()):
            if isinstance(v, torch.Tensor):
                inst.set(k, v.pin_memory())
            else:
                inst.set(k, uwcv.Boxes(v.tensor.pin_memory()))
    eng = api.Engine.get(dev)
    db = DeviceBatch(eng, batch, rank, world, dev)
    n = db.n
    counts = None
    if world > 1:
        c = torch.tensor([n], dtype=torch.int64, device=dev)
        cs = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cs, c)
        counts = cs.cpu().tolist()
    total_instances = sum(counts) if counts else n
    main = torch.cuda.current_stream(dev)
    fused = eng.fused_gather() if world > 1 else None

    def make_step(d, cts):
        def step():
            if world > 1 and fused is not None:
                # the one collective of the path, fused: the trace kernel stores the rows into every
                # rank's table over NVLink, a symmetric-memory barrier completes them
                g, gset, _total = fused.begin(cts)
                d.step(gather=g, after=lambda: fused.barrier(gset))
            elif world > 1:
                k = d.tick & 1
                ri, rf = d.rows[k]
                d.step(after=lambda: udist.all_gather_table(ri, rf, counts=cts))
            else:
                d.step()
        return step

    step = make_step(db, counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    db.check_status()
    gather_check = None
    if world > 1 and fused is not None:
        # one-time check of the fused gather against the NCCL all-gather of the same rows
        step()
        main.wait_stream(eng.trace_stream)
        main.wait_stream(eng.fill_stream)
        torch.cuda.synchronize(dev)
        k = (db.tick - 1) & 1
        ti, tf = fused.tables(fused.parity ^ 1, total_instances)
        ni, nf = udist.all_gather_table(db.rows[k][0], db.rows[k][1], counts=counts)
        same = torch.equal(ti, ni) and torch.equal(tf.nan_to_num(), nf.nan_to_num())
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_check = bool(flag.item())
        assert gather_check, "fused gather differs from the NCCL all-gather"
        barrier()

    # ---- timed region: exactly K steps, CUDA events, max over ranks -------------------
    launches0 = eng.launches
    eng.fill_events = []                      # (start, end) of every plane fill, on its own stream
    ms, wall0, wall1 = timed_steps(db, args.steps, 0, barrier, step)
    launches = eng.launches - launches0
    torch.cuda.synchronize(dev)
    fill_live = [a.elapsed_time(b) for a, b in eng.fill_events]
    eng.fill_events = None
    clocks = None
    if rank == 0:
        time.sleep(0.12)
        fields = sampler.window(wall0, wall1)
        window = "timed region"
        if len(fields) < 3:
            # the timed region is shorter than a few nvidia-smi periods: keep the same load
            # running (untimed) for about a second and sample the clocks under it
            # (rank 0 only: the kernels without the collective, which the other ranks do not join)
            c0 = time.time()
            while time.time() - c0 < 1.0:
                db.step()
                torch.cuda.synchronize(dev)
            time.sleep(0.12)
            fields = sampler.window(wall0, time.time())
            window = "timed region + 1 s continuation of the same steps"
        sampler.stop()
        clocks = sampler.summary(fields, window)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = args.steps * total_instances / (ms * 1e-3)

    # ---- per-kernel times (live CUDA events on the launching stream) -------------------
    # (each kernel group ALONE on the GPU, in call order so that the trace finds fresh marks: 1 layout, 2 | 16 tile kernel = paste without planes,
    #  8 | 16 plane fill from the tiles, 4 border trace + descriptors; 2 = the fused paste kernel of
    #  the single-stream C call, for comparison)
    reps = max(3, min(args.steps, 10))
    kt = {1: [], 2 | 16: [], 8 | 16: [], 4: [], 2: []}
    for _ in range(reps):
        for st in (1, 2 | 16, 8 | 16, 4, 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            db.stage(st)
            b.record()
            b.synchronize()
            kt[st].append(a.elapsed_time(b))
    k_layout, k_tile, k_fill, k_contour, k_paste = (statistics.mean(kt[s]) for s in (1, 2 | 16, 8 | 16, 4, 2))
    k_fill_live = statistics.mean(fill_live) if fill_live else None
    # context for the roofline: what a plain device memset of the same plane buffer reaches
    # (a write-only stream; the measured peak in MEASURED_PEAKS.json is a read+write copy)
    mt = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        db.planes.zero_()
        b.record()
        b.synchronize()
        mt.append(a.elapsed_time(b))
    memset_gbs = db.planes.numel() * 4 / (min(mt) * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    bpi = algorithmic_bytes_per_instance(H, W)
    # the dominant kernel is the plane fill; its launches inside the timed region ran NEXT TO the
    # border trace of the same step and the tile kernel of the following one (that is the point of
    # the split pipeline), so its live duration there is the honest denominator; alone it is faster
    bpi_fill = H * W // 8                       # the plane bytes it writes (tile words read: ~0.7 KB more)
    achieved = n * bpi_fill / ((k_fill_live or k_fill) * 1e-3) / 1e9
    achieved_alone = n * bpi_fill / (k_fill * 1e-3) / 1e9
    traffic, traffic_source = None, None
    tp = os.path.join(ROOT, "profiles", "plane_fill_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_instance"] * n
            traffic_source = ("replayed from the committed ncu --set full capture, not measured in this "
                              "run: " + tj.get("source", "profiles/plane_fill_traffic.json"))
        except Exception:
            traffic = None

    # ---- rows-only contract (no full-frame planes): the arithmetic alone, compute-bound --------
    barrier()
    ro_steps = max(3, min(args.steps, 20))
    ro_ms, _, _ = timed_steps(db, ro_steps, 3, barrier, lambda: db.step(planes=False))
    if world > 1:
        t = torch.tensor([ro_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ro_ms = float(t.item())
    rows_only = {"ms_per_step": ro_ms / ro_steps, "value": ro_steps * total_instances / (ro_ms * 1e-3),
                 "unit": UNIT, "steps": ro_steps,
                 "note": "same step without writing the full-frame planes (cropped contract, SURVEY 8(d)): "
                         "~4.2 KB algorithmic bytes per instance, bound by the paste arithmetic and the "
                         "serial border walks, not by HBM; no gather in this leg"}

    # ---- weak-scaling line beside the strong configs[2] one (64 images per GPU) -----------------
    weak = None
    if strong:
        if n_img == IMAGES_PER_GPU:
            weak = {"value": value, "ms_per_step": ms / args.steps, "images_per_gpu": n_img,
                    "note": "identical to the main line at this N"}
        else:
            planes_main = db.planes
            db.planes = None
            del planes_main
            wb = batch[:IMAGES_PER_GPU]
            if len(wb) < IMAGES_PER_GPU:             # N = 8: the rank's shard holds only 32 images
                wb = wb + synth.blob_batch(IMAGES_PER_GPU - len(wb), n_inst, H, W, seed=1234,
                                           first_image=rank + len(wb) * world, stride=world)
            wdb = DeviceBatch(eng, wb, rank, world, dev)
            wc = torch.tensor([wdb.n], dtype=torch.int64, device=dev)
            wcs = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(wcs, wc)
            wcounts = wcs.cpu().tolist()
            wsteps = max(3, min(args.steps, 20))
            wms, _, _ = timed_steps(wdb, wsteps, 3, barrier, make_step(wdb, wcounts))
            t = torch.tensor([wms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wms = float(t.item())
            weak = {"value": wsteps * sum(wcounts) / (wms * 1e-3), "ms_per_step": wms / wsteps,
                    "images_per_gpu": IMAGES_PER_GPU, "steps": wsteps, "unit": UNIT}
            del wdb

    # ---- e2e: the public call with pinned host inputs -----------------------------------
    h2d = sum(int(b.pred_masks.numel()) * 4 + len(b) * (16 + 4 + 8 + 4 + 4) for b in batch)
    d2h = n * (20 * 8 + 30 * 8) + 32 + (8 if world > 1 else 0)   # every rank: its own rows (+ flag word)
    db.planes = None
    del db
    # at N > 1 the whole job's table is gathered to the host of rank 0: every rank copies its own
    # rows into ONE host table shared by the ranks of the node (uwcv.dist.SharedHostTable); nothing
    # moves between the devices and no rank reads another rank's rows over PCIe
    kw = dict(write_planes=True, gather=world > 1, gather_counts=counts,
              gather_dst=0 if world > 1 else None)
    sink = None
    if world > 1:
        sink = "shared host table" if eng.host_table() is not None else "device all-gather + one read"
    if os.environ.get("UWCV_BENCH_E2E_NOGATHER"):            # probe: independent ranks
        kw = dict(write_planes=True)
    e2e_steps = args.steps
    for _ in range(3):
        table = uwcv.measure_instances(batch, (H, W), device=dev, **kw)
    # (a) one synchronous call per step (latency form)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        table = uwcv.measure_instances(batch, (H, W), device=dev, **kw)
    barrier()
    sync_s = time.perf_counter() - t0
    del table
    # (b) the throughput form of the same call: uwcv.MeasurementStream keeps two calls in
    #     flight, so the H2D of step i + 1 and the D2H of step i - 1 run under the kernels of
    #     step i.  Every step still copies its inputs from pinned host memory and reads its
    #     rows back; all K tables are materialised on the host inside the timed region.
    stream = uwcv.MeasurementStream(dev, depth=args.e2e_depth)
    for table in stream.map((batch for _ in range(stream.depth + 3)), (H, W), **kw):
        pass
    barrier()
    t0 = time.perf_counter()
    got = 0
    crc = 0
    for table in stream.map((batch for _ in range(e2e_steps)), (H, W), **kw):
        got += len(table)
        crc ^= int(table.ints[-1, 5]) if len(table) else 0      # touch the last row on the host
    barrier()
    e2e_s = time.perf_counter() - t0
    del table
    if not os.environ.get("UWCV_BENCH_E2E_NOGATHER"):
        assert got == e2e_steps * (total_instances if rank == 0 else n), (got, total_instances, n)
    if os.environ.get("UWCV_BENCH_VERBOSE"):
        print(f"rank {rank}: e2e {e2e_s / e2e_steps * 1e3:.2f} ms/step, sync {sync_s / e2e_steps * 1e3:.2f}",
              file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([e2e_s, sync_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, sync_s = float(t[0].item()), float(t[1].item())
    e2e_val = e2e_steps * total_instances / e2e_s

    # ---- (c) the same call with DEVICE-resident inputs (what a caller holding the network's
    #      outputs on the GPU pays: no mask H2D, rows still read back to the host every step)
    #      The mask probabilities of the batch sit back to back in ONE device allocation, split per
    #      image, as a mask head's output does: the call uses them where they are.)
    dbatch = []
    all_masks = torch.cat([inst.pred_masks for inst in batch]).to(dev)
    lo = 0
    for inst in batch:
        o = uwcv.Instances(inst.image_size)
        for k, v in inst.get_fields().items():
            if k == "pred_masks":
                o.set(k, all_masks[lo:lo + len(inst)])
            else:
                o.set(k, uwcv.Boxes(v.tensor.to(dev)) if hasattr(v, "tensor") else v.to(dev))
        lo += len(inst)
        dbatch.append(o)
    for table in stream.map((dbatch for _ in range(3)), (H, W), **kw):
        pass
    barrier()
    t0 = time.perf_counter()
    for table in stream.map((dbatch for _ in range(e2e_steps)), (H, W), **kw):
        pass
    barrier()
    dres_s = time.perf_counter() - t0
    del table, dbatch, all_masks
    if world > 1:
        t = torch.tensor([dres_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dres_s = float(t.item())

    # ---- other configs + CPU baselines (rank 0, N = 1 only) ---------------------------------
    cpu = cpu_b = others = lib_gpu = None
    if world == 1 and not args.no_other_configs:
        others = other_configs(eng, dev)
    if world == 1 and not args.no_library_baseline:
        try:
            lib_gpu = library_baseline_gpu(eng, dev)
            torch.cuda.empty_cache()
        except Exception as e:                       # a baseline leg never takes the bench line down
            lib_gpu = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_cpu_baseline:
        nc = 600                 # ~13 s of CPU work for baseline A on 16 cores
        sample = synth.blob_instances(0, nc, H, W, seed=1234)
        t0 = time.perf_counter()
        done, _ = reference_step(sample, (H, W))
        dt = time.perf_counter() - t0
        cpu = {"value": done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 image 2048x2048 x {nc} instances, detector_postprocess + "
                         f"4 x GetMask_Contours (oracle restatement), {dt:.1f} s"}
        # BASELINE.md section 4, baseline B ("per-instance"): oracle paste -> cv2.moments, bbox,
        # external contours, descriptor block per instance
        from oracle import pipeline as P
        t0 = time.perf_counter()
        ri_, _rf = P.oracle_table([sample], (H, W))
        dt = time.perf_counter() - t0
        cpu_b = {"value": ri_.shape[0] / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                 "sample": f"1 image 2048x2048 x {nc} instances, per-instance CPU paste + cv2.moments + "
                           f"findContours + descriptor block (oracle.pipeline.oracle_table), {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD2 if strong else WORKLOAD, "images_per_gpu": n_img,
                       "images_total": n_img * world,
                       "instances_per_image": n_inst, "image": f"{H}x{W}",
                       "instances_per_gpu": n, "mask_output": "full-frame bit-planes in HBM",
                       "plane_memory": ("compressible device memory (CUDA VMM, generic compression: the planes "
                                        "are ~97 % zero words, which B200 compresses between L2 and HBM; "
                                        "uwcv_planes_alloc)" if eng.compressible_planes
                                        else "ordinary device memory (compression not granted)"),
                       "l2": "inputs (200 MB) and outputs (33.5 GB) per 64 images exceed the 126 MB L2",
                       "collective": "none" if world == 1 else
                       ("all-gather of the row table fused into the trace kernel (peer stores "
                        "over NVLink into symmetric memory + signal barrier)" if fused is not None
                        else "NCCL all_gather of the row table")},
            "mp_per_sec": args.steps * world * n_img * H * W / 1e6 / (ms * 1e-3),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "call": f"uwcv.MeasurementStream(depth={args.e2e_depth}).map (pinned host Instances in, "
                            "host MeasurementTable out, every step)",
                    "sync_call_ms_per_step": sync_s / e2e_steps * 1e3,
                    "h2d_gbs_whole_job": h2d * world / (e2e_s / e2e_steps) / 1e9,
                    "device_resident_inputs_ms_per_step": dres_s / e2e_steps * 1e3,
                    "device_resident_inputs_value": e2e_steps * total_instances / dres_s},
            "gpu_launches": launches,
            "kernel_ms": {"layout": k_layout, "tile_measure": k_tile, "plane_fill": k_fill,
                          "plane_fill_in_pipeline": k_fill_live, "contour": k_contour,
                          "fused_paste_measure_single_stream": k_paste,
                          "note": "each kernel group alone on the GPU (CUDA events), except "
                                  "plane_fill_in_pipeline: mean over the timed region's launches, next to "
                                  "the trace of the same step and the tile kernel of the next one"},
            "pipeline": "split: layout + tile kernel (main stream) -> [plane fill (fill stream) || border "
                        "trace (trace stream)], two workspaces in turn; uwcv.Engine.run_overlapped",
            "rows_only": rows_only,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                         "peak_source": peak_src,
                         "kernel": "plane_fill_kernel",
                         "compressed_planes": bool(eng.compressible_planes),
                         "note": "achieved = plane bytes DELIVERED per second; in compressible memory fewer "
                                 "bytes reach HBM (see traffic), so the fraction of the HBM copy peak can pass 1",
                         "timed": "mean duration of its launches inside the timed region (CUDA events on "
                                  "the fill stream), i.e. while the trace / tile kernels share the GPU",
                         "achieved_alone": achieved_alone, "frac_alone": achieved_alone / peak,
                         "bytes_per_instance": bpi_fill, "instances_per_launch": n,
                         "bytes_per_instance_whole_path": bpi,
                         "whole_step_gbs": n * bpi / (ms / args.steps * 1e-3) / 1e9,
                         "frac_of_nominal_8000": achieved / 8000.0,
                         # a plain zero fill of the same plane buffer (torch's one-shot
                         # elementwise kernel): the ceiling of a write-only stream on this GPU
                         "memset_same_buffer_gbs": memset_gbs,
                         "frac_of_zero_fill": achieved / memset_gbs},
            "clocks": clocks,
        }
        if world > 1:
            line["e2e"]["gather_sink"] = sink
            line["e2e"]["d2h_note"] = "per rank: its own rows into the node-shared host table"
        if gather_check is not None:
            line["config"]["fused_gather_equals_nccl"] = gather_check
        if weak is not None:
            line["weak_scaling"] = weak
        if others is not None:
            line["other_configs"] = others
        if lib_gpu is not None:
            line["library_baseline_gpu"] = lib_gpu
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if cpu_b is not None:
            line["cpu_baseline_b"] = cpu_b
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200,
                    help="timed steps (default: a timed region above one second)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=None,
                    help="images per GPU (default: 64 at N = 1; 256 / N at N > 1, configs[2])")
    ap.add_argument("--e2e-depth", type=int, default=3,
                    help="calls in flight of the streamed e2e leg (uwcv.MeasurementStream)")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--instances", type=int, default=INSTANCES_PER_IMAGE)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip the torch / torchvision CUDA baseline of the same GPU (library_baseline_gpu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
