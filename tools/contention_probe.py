"""Which co-runner stretches the plane fill of the split pipeline?  The fill (stage 8) of one workspace is
timed alone, next to the border trace of the same workspace, next to the tile kernel of the other
workspace, and next to both.

    python tools/contention_probe.py [--variant tuning] [--images 64]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="")
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--instances", type=int, default=1000)
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
from uwcv import _lib  # noqa: E402
if a.variant:
    _lib.use_library_variant(a.variant)
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
planes = eng.alloc_planes(n, H, W)
ri = torch.empty((n, 20), dtype=torch.int64, device=dev)
rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
ri2, rf2 = ri.clone(), rf.clone()
kw = dict(classes=d_classes, scores=d_scores, planes=planes, n_tile_words=words)
sa, sb, sc = torch.cuda.Stream(dev, priority=-2), torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev)


def run(st, slot, rows=(ri, rf)):
    eng.run(d_masks, d_boxes, H, W, rows_i=rows[0], rows_f=rows[1], stages=st, ws_slot=slot, **kw)


def case(with_trace, with_tile):
    out = {"fill": [], "trace": [], "tile": []}
    for _ in range(a.reps):
        run(1 | 2 | 16, 0)                 # tiles of workspace 0, marks cleared
        run(1, 1, (ri2, rf2))              # layout of workspace 1
        torch.cuda.synchronize()
        ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in out}
        with torch.cuda.stream(sa):
            ev["fill"][0].record(); run(8 | 16, 0); ev["fill"][1].record()
        if with_trace:
            with torch.cuda.stream(sb):
                ev["trace"][0].record(); run(4 | 16, 0); ev["trace"][1].record()
        if with_tile:
            with torch.cuda.stream(sc):
                ev["tile"][0].record(); run(2 | 16, 1, (ri2, rf2)); ev["tile"][1].record()
        torch.cuda.synchronize()
        out["fill"].append(ev["fill"][0].elapsed_time(ev["fill"][1]))
        if with_trace:
            out["trace"].append(ev["trace"][0].elapsed_time(ev["trace"][1]))
        if with_tile:
            out["tile"].append(ev["tile"][0].elapsed_time(ev["tile"][1]))
    return {k: round(min(v), 3) for k, v in out.items() if v}


res = {"n": n, "knobs": {k: v for k, v in os.environ.items() if k.startswith("UWCV_")}}
run(7, 0); run(7, 1, (ri2, rf2)); torch.cuda.synchronize()
res["fill_alone"] = case(False, False)
res["fill_trace"] = case(True, False)
res["fill_tile"] = case(False, True)
res["fill_trace_tile"] = case(True, True)
print(json.dumps(res))
