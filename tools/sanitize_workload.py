import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/uw-com-vision_b200")
import torch, numpy as np, uwcv
from uwcv import synth, api
from oracle import pipeline as P
H, W = 200, 333
batch = [synth.blob_instances(k, 30, H, W, seed=11, size_range=(2.0, 150.0)) for k in range(2)]
t, planes = uwcv.measure_instances(batch, (H, W), return_planes=True)
ri, rf = P.oracle_table(batch, (H, W))
assert np.array_equal(t.ints, ri)
m = uwcv.paste_masks_in_image(batch[0].pred_masks[:, 0], batch[0].pred_boxes.tensor, (H, W))
b, s, c = synth.clustered_candidates(300, H, W, seed=3, n_clusters=4)
eng = api.Engine.get()
keep, cnt = eng.nms(b.cuda(), s.cuda(), c.cuda(), [0, 500, len(b)], 0.05, 0.5, 100)
torch.cuda.synchronize()
print("sanitize workload ok", len(t), int(cnt.sum()))
