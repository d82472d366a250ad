import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if '__file__' in dir() else '.'
sys.path.insert(0, '.'); sys.path.insert(0, 'uw-com-vision_b200')
import torch, uwcv
from uwcv import api, synth
H = W = 2048
dev = torch.device("cuda", 0)
eng = api.Engine.get(dev)
batch = synth.blob_batch(64, 1000, H, W, seed=1234)
boxes = torch.cat([api.scale_clip_boxes(b.pred_boxes.tensor, (H, W), (H, W))[0] for b in batch])
n = len(boxes); words = api.tile_words(boxes, H, W)
d_boxes = boxes.to(dev); d_masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous().to(dev)
ri = torch.empty((n, 20), dtype=torch.int64, device=dev); rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
for _ in range(4):
    eng.run(d_masks, d_boxes, H, W, n_tile_words=words, rows_i=ri, rows_f=rf)
torch.cuda.synchronize()
