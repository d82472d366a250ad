"""Randomised parity sweep of the union / connected-component mode (reference-literal
GetMask_Contours rows) against the CPU oracle.  python tools/union_sweep.py [n_cases] [seed0]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import numpy as np, torch
import uwcv
from uwcv import synth
from oracle import pipeline as P

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 900
tot = dict(cases=0, rows=0, shape_mismatch=0, rows_over_1e6=0, max_rel=0.0)
t0 = time.time()
for case in range(n_cases):
    g = torch.Generator().manual_seed(seed0 + case)
    H = int(torch.randint(96, 520, (1,), generator=g)); W = int(torch.randint(96, 520, (1,), generator=g))
    n = int(torch.randint(5, 140, (1,), generator=g))
    hi = float(20 + torch.rand(1, generator=g) * 120)
    inst = synth.blob_instances(case, n, H, W, seed=seed0 + case, size_range=(8.0, hi))
    for cls in ([0], [1], [2], [3], [0, 1, 2, 3]):
        ut = uwcv.measure_union(inst, (H, W), classes_of_interest=cls)
        try:
            ref = P.reference_literal_rows(inst, (H, W), cls)
        except ValueError:
            ref = None
        ref = np.zeros((0, 9)) if ref is None else ref
        mine = ut.reference_rows()
        tot["rows"] += len(ref)
        if mine.shape != ref.shape:
            tot["shape_mismatch"] += 1
            continue
        if len(ref):
            rel = np.abs(mine - ref) / np.maximum(np.abs(ref), 1e-30)
            tot["rows_over_1e6"] += int((rel.max(axis=1) > 1e-6).sum())
            tot["max_rel"] = max(tot["max_rel"], float(rel.max()))
    tot["cases"] += 1
tot["seconds"] = round(time.time() - t0, 1); tot["seed0"] = seed0
print(json.dumps(tot))
