"""Wall-time split of uwcv.measure_union on 8 configs[1]-style images (class 3)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
import uwcv
from uwcv import synth
import cProfile, pstats
batch = synth.blob_batch(8, 1000, 2048, 2048, seed=1234)
for _ in range(2):
    ut = uwcv.measure_union(batch, (2048, 2048), classes_of_interest=[3])
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); ut = uwcv.measure_union(batch, (2048, 2048), classes_of_interest=[3]); ts.append((time.perf_counter() - t0) * 1e3)
print("wall ms", [round(t, 2) for t in ts], "rows", len(ut))
pr = cProfile.Profile(); pr.enable()
ut = uwcv.measure_union(batch, (2048, 2048), classes_of_interest=[3])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
