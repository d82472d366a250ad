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
