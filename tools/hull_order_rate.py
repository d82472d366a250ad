"""CPU probe (oracle only, no GPU): how often the device's hull-order rule ("the contour's start pixel comes last",
oracle.contours_py.hull_like_cv) and OpenCV's (convex_hull_cv) give different min-area rectangles on the bench
workload -- configs[1] blobs, best external contour of every instance.  Usage: hull_order_rate.py [images of 1000 instances]"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'uw-com-vision_b200'))
import numpy as np, torch, cv2
from uwcv import synth
from oracle import pipeline as P, contours_py as cp, measure as M
H = W = 2048
t0 = time.time()
tot = nonsimple = order_diff = rect_diff = 0
for b in synth.blob_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 20, 1000, H, W, seed=1234):
    res = P.postprocess_boxes(P.to_oracle_instances(b), (H, W))
    for win, y0, x0 in P.oracle_windows(res):
        m = win.astype(np.uint8)
        if not m.any(): continue
        cnts, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE, offset=(x0, y0))
        best, _n = M.pick_best_contour(cnts)
        p = best.reshape(-1, 2)
        tot += 1
        simple = len({tuple(q) for q in p.tolist()}) == len(p)
        nonsimple += not simple
        if simple: continue
        hd = cp.hull_like_cv(p); hc = cp.convex_hull_cv(p)
        if not np.array_equal(hd, hc):
            order_diff += 1
            if cp.min_area_rect_cv(hd) != cp.min_area_rect_cv(hc): rect_diff += 1
print(dict(instances=tot, non_simple=nonsimple, hull_start_differs=order_diff, rect_differs=rect_diff, seconds=round(time.time()-t0,1)))
