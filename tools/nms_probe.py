import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
from uwcv import api, synth
dev = torch.device("cuda", 0)
eng = api.Engine.get(dev)
cb, cs, cc = synth.clustered_candidates(5000, 4096, 4096, seed=99)
dcb, dcs, dcc = cb.to(dev), cs.to(dev), cc.to(dev)
for _ in range(3):
    keep, cnt = eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    keep, cnt = eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000)
b.record(); b.synchronize()
print("nms ms", a.elapsed_time(b) / 5, "kept", int(cnt[0]))
# Detectron2-scale: 8 images x 1000 proposals x 80 classes, threshold 0.8 keeps a few hundred candidates
import ctypes as C
g = torch.Generator().manual_seed(3)
R, K, B = 1000, 80, 8
n = R * K
xy = torch.rand(B * n, 2, generator=g) * 900
wh = torch.rand(B * n, 2, generator=g) * 100 + 4
bx = torch.cat([xy, xy + wh], dim=1).to(dev)
sc = torch.rand(B * n, generator=g).pow(8).to(dev)          # few scores above 0.8
cl = torch.arange(K).repeat(B * R).to(dev)
off = [i * n for i in range(B + 1)]
for _ in range(3):
    keep, cnt = eng.nms(bx, sc, cl, off, 0.8, 0.5, 100, num_classes=K)
torch.cuda.synchronize()
a.record()
for _ in range(5):
    keep, cnt = eng.nms(bx, sc, cl, off, 0.8, 0.5, 100, num_classes=K)
b.record(); b.synchronize()
offs = (C.c_int64 * (B + 1))(*off)
print("d2-scale: 8 x 80000 candidates, nms ms", a.elapsed_time(b) / 5, "kept", cnt.tolist(),
      "workspace MB", eng.L.uwcv_nms_workspace_bytes(offs, B, K) / 1e6)
