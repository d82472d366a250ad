"""Probe: border trace of step i on a second stream under the paste kernel of step i + 1
(double-buffered workspace and row tables) vs the serial step.  Env knobs: UWCV_FILL=2 (evict-first
zero rows), UWCV_PASTE_CTAS -- read by the tuning library only: python tools/overlap_probe.py tuning."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
from uwcv import _lib
if len(sys.argv) > 1:
    _lib.use_library_variant(sys.argv[1])
import uwcv
from uwcv import api, synth
H = W = 2048
dev = torch.device("cuda", 0)
eng = api.Engine.get(dev)
batch = synth.blob_batch(64, 1000, H, W, seed=1234)
boxes = torch.cat([api.scale_clip_boxes(b.pred_boxes.tensor, (H, W), (H, W))[0] for b in batch])
n = len(boxes); words = api.tile_words(boxes, H, W)
d_boxes = boxes.to(dev); d_masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous().to(dev)
d_scores = torch.cat([b.scores for b in batch]).to(dev); d_classes = torch.cat([b.pred_classes for b in batch]).to(dev)
planes = eng.alloc_planes(n, H, W)
rows = [(torch.empty((n, 20), dtype=torch.int64, device=dev), torch.empty((n, 30), dtype=torch.float64, device=dev)) for _ in range(2)]
status = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(2)]
common = dict(classes=d_classes, scores=d_scores, planes=planes, n_tile_words=words)
main = torch.cuda.current_stream(dev); side = torch.cuda.Stream(dev)
ev_p = [torch.cuda.Event() for _ in range(2)]; ev_c = [torch.cuda.Event() for _ in range(2)]

def serial(k):
    ri, rf = rows[0]
    eng.run(d_masks, d_boxes, H, W, rows_i=ri, rows_f=rf, stages=7, **common)

def overlapped(k):
    p = k & 1
    ri, rf = rows[p]
    main.wait_event(ev_c[p])
    eng.run(d_masks, d_boxes, H, W, rows_i=ri, rows_f=rf, stages=3, ws_slot=p, status=status[p], **common)
    ev_p[p].record(main)
    with torch.cuda.stream(side):
        side.wait_event(ev_p[p])
        eng.run(d_masks, d_boxes, H, W, rows_i=ri, rows_f=rf, stages=4, ws_slot=p, status=status[p], **common)
        ev_c[p].record(side)

def make_intra(chunks):
    bounds = [(n * c // chunks, n * (c + 1) // chunks) for c in range(chunks)]
    evs = [torch.cuda.Event() for _ in range(chunks)]
    def intra(k):
        ri, rf = rows[0]
        main.wait_stream(side)                 # the previous step's traces read this workspace
        eng.run(d_masks, d_boxes, H, W, rows_i=ri, rows_f=rf, stages=1, **common)
        for c, (lo, hi) in enumerate(bounds):
            eng.run(d_masks, d_boxes, H, W, rows_i=ri, rows_f=rf, stages=2, first=lo, count=hi - lo, **common)
            evs[c].record(main)
            with torch.cuda.stream(side):
                side.wait_event(evs[c])
                eng.run(d_masks, d_boxes, H, W, rows_i=ri, rows_f=rf, stages=4, first=lo, count=hi - lo, **common)
    return intra


def timed(fn, K=10):
    for k in range(4): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for k in range(K): fn(k)
    main.wait_event(ev_c[0]); main.wait_event(ev_c[1]); main.wait_stream(side)
    b.record(main); b.synchronize()
    return a.elapsed_time(b) / K

eng._workspace(n, words, 0); eng._workspace(n, words, 1)
for e in ev_c: e.record(main)
print("knobs", {k: v for k, v in os.environ.items() if k.startswith("UWCV_")})
print("serial   ms/step", round(timed(serial), 3))
print("overlap  ms/step", round(timed(overlapped), 3))
for ch in (2, 4, 8):
    print(f"intra-step chunks {ch} ms/step", round(timed(make_intra(ch)), 3))
ref = rows[0][0].clone(), rows[0][1].clone()
serial(0); torch.cuda.synchronize()
assert torch.equal(ref[0], rows[0][0]) and torch.equal(ref[1].nan_to_num(), rows[0][1].nan_to_num()), "rows differ"
print("rows identical")
