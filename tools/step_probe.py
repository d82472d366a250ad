"""Device-resident step timing of configs[1] (or a smaller batch): kernels alone, the serial step and
the overlapped step (Engine.run_overlapped), without the e2e / CPU legs of bench.py.

    python tools/step_probe.py [--variant tuning] [--images 64] [--steps 10]

With --variant tuning the -DUWCV_TUNING library is loaded and UWCV_* environment knobs apply."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="")
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--instances", type=int, default=1000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--no-planes", action="store_true")
a = ap.parse_args()
from uwcv import _lib  # noqa: E402
if a.variant:
    _lib.use_library_variant(a.variant)
import uwcv  # noqa: E402
from uwcv import api, synth  # noqa: E402

H = W = a.size
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
batch = synth.blob_batch(a.images, a.instances, H, W, seed=1234)
eng = api.Engine.get(dev)
boxes = torch.cat([api.scale_clip_boxes(b.pred_boxes.tensor, (H, W), (H, W))[0] for b in batch])
n = int(boxes.shape[0])
words = api.tile_words(boxes, H, W)
d_boxes = boxes.to(dev)
d_masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous().to(dev)
d_scores = torch.cat([b.scores for b in batch]).to(dev)
d_classes = torch.cat([b.pred_classes for b in batch]).to(dev)
d_img = torch.cat([torch.full((len(b),), i, dtype=torch.int32) for i, b in enumerate(batch)]).to(dev)
d_inst = torch.cat([torch.arange(len(b), dtype=torch.int32) for b in batch]).to(dev)
planes = None if a.no_planes else eng.alloc_planes(n, H, W)
rows = [(torch.empty((n, 20), dtype=torch.int64, device=dev),
         torch.empty((n, 30), dtype=torch.float64, device=dev)) for _ in range(2)]
stat = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(2)]
kw = dict(image_idx=d_img, inst_idx=d_inst, classes=d_classes, scores=d_scores, planes=planes,
          n_tile_words=words)
main = torch.cuda.current_stream(dev)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        out.append(e0.elapsed_time(e1))
    return statistics.mean(out), min(out)


res = {"n": n, "knobs": {k: v for k, v in os.environ.items() if k.startswith("UWCV_")}}
for _ in range(3):
    eng.run(d_masks, d_boxes, H, W, rows_i=rows[0][0], rows_f=rows[0][1], **kw)
torch.cuda.synchronize()
assert int(eng.status.cpu()[0]) == 0
for name, st in (("layout", 1), ("paste", 2), ("trace", 4), ("serial_step", 7)):
    res[name] = timed(lambda: eng.run(d_masks, d_boxes, H, W, rows_i=rows[0][0], rows_f=rows[0][1],
                                      stages=st, **kw), max(3, a.steps))
tick = [0]


def ostep():
    k = tick[0] & 1
    tick[0] += 1
    eng.run_overlapped(d_masks, d_boxes, H, W, rows_i=rows[k][0], rows_f=rows[k][1], status=stat[k], **kw)


for _ in range(3):
    ostep()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    ostep()
main.wait_stream(eng.trace_stream)
e1.record(); e1.synchronize()
res["overlapped_step"] = e0.elapsed_time(e1) / a.steps
if a.variant.startswith("tuning"):
    import ctypes as C
    L = _lib.lib()
    buf = (C.c_ulonglong * 4)()
    L.uwcv_tuning_fused_stats(buf, 1)
    eng.run(d_masks, d_boxes, H, W, rows_i=rows[0][0], rows_f=rows[0][1], **kw)
    torch.cuda.synchronize()
    L.uwcv_tuning_fused_stats(buf, 1)
    res["fused_stats_one_step"] = dict(alloc_retries=buf[0], tracer_idle_polls=buf[1], tracer_warp_iters=buf[2],
                                       tracer_lane_steps=buf[3])
res["checksum"] = [int(rows[0][0].sum().item()), float(rows[0][1].nan_to_num().sum().item())]
print(json.dumps(res))
