"""Short target for ncu: configs[1]-style batch (default 16 images = 16 000 instances, full-frame planes),
3 warm steps + 1 step through Engine.run, every kernel back to back on one stream.  Second argument "split"
(default): the stages of the split pipeline (layout, tile kernel, plane fill, trace); "fused": stages = 7."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
from uwcv import api, synth
images = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mode = sys.argv[2] if len(sys.argv) > 2 else "split"
stages = (1 | 2 | 16 | 8 | 4) if mode == "split" else 7
H = W = 2048
dev = torch.device("cuda", 0)
eng = api.Engine.get(dev)
batch = synth.blob_batch(images, 1000, H, W, seed=1234)
boxes = torch.cat([api.scale_clip_boxes(b.pred_boxes.tensor, (H, W), (H, W))[0] for b in batch])
n = len(boxes); words = api.tile_words(boxes, H, W)
d_boxes = boxes.to(dev); d_masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous().to(dev)
d_scores = torch.cat([b.scores for b in batch]).to(dev)
planes = eng.alloc_planes(n, H, W)
ri = torch.empty((n, 20), dtype=torch.int64, device=dev); rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
for _ in range(4):
    eng.run(d_masks, d_boxes, H, W, planes=planes, scores=d_scores, n_tile_words=words, rows_i=ri, rows_f=rf,
            stages=stages)
torch.cuda.synchronize()
assert int(eng.status.cpu()[0]) == 0
print("ok", n, int(ri[:, 5].sum()))
