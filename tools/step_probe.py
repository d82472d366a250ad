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
ap.add_argument("--check-split", action="store_true", help="planes / rows of the split pipeline == fused kernel")
ap.add_argument("--timeline", default="", help="write the CUPTI kernel timeline of 6 split steps to this file")
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
if planes is not None:
    for name, st in (("tile_rows_only", 2 | 16), ("plane_fill", 8 | 16), ("serial_split_step", 1 | 2 | 16 | 8 | 4)):
        res[name] = timed(lambda: eng.run(d_masks, d_boxes, H, W, rows_i=rows[0][0], rows_f=rows[0][1],
                                          stages=st, **kw), max(3, a.steps))
tick = [0]


def ostep(split):
    k = tick[0] & 1
    tick[0] += 1
    eng.run_overlapped(d_masks, d_boxes, H, W, rows_i=rows[k][0], rows_f=rows[k][1], status=stat[k],
                       split=split, **kw)


for split in ((False, True) if planes is not None else (False,)):
    for _ in range(3):
        ostep(split)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ostep(split)
    main.wait_stream(eng.trace_stream)
    main.wait_stream(eng.fill_stream)
    e1.record(); e1.synchronize()
    res["overlapped_split_step" if split else "overlapped_step"] = e0.elapsed_time(e1) / a.steps
    res["rows_checksum_split" if split else "rows_checksum_fused"] = [
        int(rows[0][0].sum().item()), float(rows[0][1].nan_to_num().sum().item()),
        int(rows[1][0].sum().item()), float(rows[1][1].nan_to_num().sum().item())]
if a.check_split and planes is not None:
    ref_i, ref_f = rows[0][0].clone(), rows[0][1].clone()
    eng.run(d_masks, d_boxes, H, W, rows_i=ref_i, rows_f=ref_f, **kw)          # fused kernel
    torch.cuda.synchronize()
    ref_planes = planes.clone()
    planes.fill_(-1)
    ostep(True); ostep(True)
    torch.cuda.synchronize()
    res["split_planes_equal"] = bool(torch.equal(ref_planes, planes))
    res["split_rows_equal"] = bool(torch.equal(ref_i, rows[0][0]) and torch.equal(ref_i, rows[1][0]) and
                                   torch.equal(ref_f.nan_to_num(), rows[0][1].nan_to_num()) and
                                   torch.equal(ref_f.nan_to_num(), rows[1][1].nan_to_num()))
    del ref_planes
if a.timeline and planes is not None:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(6):
            ostep(True)
        torch.cuda.synchronize()
    tmp = a.timeline + ".chrome.json"
    prof.export_chrome_trace(tmp)
    ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset") and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    with open(a.timeline, "w") as f:
        for e in ev:
            f.write(f"{(e['ts'] - ev[0]['ts']) / 1e3:9.3f} ms  +{e['dur'] / 1e3:7.3f} ms  stream {e.get('args', {}).get('stream')}  "
                    f"{e['name'].split('(')[0][-44:]}\n")
    os.remove(tmp)
res["checksum"] = [int(rows[0][0].sum().item()), float(rows[0][1].nan_to_num().sum().item())]
print(json.dumps(res))
