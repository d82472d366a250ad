"""Probe: how the contour kernel time depends on the speckle fraction / contour counts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import numpy as np, torch
import uwcv
from uwcv import api, synth

H = W = 2048
dev = torch.device("cuda", 0)
eng = api.Engine.get(dev)
for frac in (0.0, 0.1):
    g = torch.Generator().manual_seed(0)
    batch = []
    for k in range(16):
        inst = synth.blob_instances(k, 1000, H, W, seed=1234)
        gg = torch.Generator().manual_seed(5000 + k)
        inst.remove("pred_masks"); inst.set("pred_masks", synth.blob_probs(len(inst), gg, speckle_frac=frac)[:, None])
        batch.append(inst)
    boxes = torch.cat([b.pred_boxes.tensor for b in batch]).to(dev)
    masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous().to(dev)
    n = len(boxes)
    words = api.tile_words(boxes, H, W)
    ri = torch.empty((n, 20), dtype=torch.int64, device=dev); rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
    def run(st): eng.run(masks, boxes, H, W, n_tile_words=words, rows_i=ri, rows_f=rf, stages=st)
    for _ in range(3): run(7)
    torch.cuda.synchronize()
    print('status', eng.status.cpu().tolist(), 'area sum', int(ri[:,5].sum()), 'ncont sum', int(ri[:,4].sum()))
    ts = {1: [], 2: [], 4: []}
    for _ in range(5):
        for st in (1, 2, 4):          # stage 2 resets the marks stage 4 consumes
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(st); b.record(); b.synchronize(); ts[st].append(a.elapsed_time(b))
    ts = {k: min(v) for k, v in ts.items()}
    nc = ri[:, 4].cpu().numpy(); npts = ri[:, 19].cpu().numpy()
    print(f"speckle {frac}: n={n} layout {ts[1]:.3f} paste(cropped) {ts[2]:.3f} contour {ts[4]:.3f} ms | "
          f"n_contours mean {nc.mean():.2f} p50 {np.percentile(nc,50)} p90 {np.percentile(nc,90)} p99 {np.percentile(nc,99)} max {nc.max()} | "
          f"tile words/inst {words/n:.1f}")
