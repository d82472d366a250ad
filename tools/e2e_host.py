"""Probe: host-side submit / result times of the streamed path, with and without planes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
from uwcv import _lib as _uwcv_lib
if os.environ.get("UWCV_TEST_VARIANT"):                  # e.g. "tuning": UWCV_* knobs apply
    _uwcv_lib.use_library_variant(os.environ["UWCV_TEST_VARIANT"])
import uwcv
from uwcv import synth
H = W = 2048
batch = synth.blob_batch(64, 1000, H, W, seed=1234)
for inst in batch:
    for k, v in list(inst.get_fields().items()):
        inst.set(k, v.pin_memory() if isinstance(v, torch.Tensor) else uwcv.Boxes(v.tensor.pin_memory()))
K = int(os.environ.get('E2E_STEPS', '12'))
ONLY = os.environ.get('E2E_ONLY')          # 'planes,depth'
CH = int(os.environ.get('E2E_CHUNKS', '4'))
for planes in (False, True):
    for depth in (2, 3):
        if ONLY and ONLY != f"{int(planes)},{depth}":
            continue
        st = uwcv.MeasurementStream(depth=depth)
        for _ in st.map((batch for _ in range(4)), (H, W), write_planes=planes, pipeline_chunks=CH): pass
        torch.cuda.synchronize()
        ts, tr, pend = [], [], []
        t_all = time.perf_counter()
        for i in range(K):
            a = time.perf_counter(); pend.append(st.submit(batch, (H, W), write_planes=planes, pipeline_chunks=CH)); b = time.perf_counter()
            ts.append(b - a)
            if len(pend) >= depth:
                a = time.perf_counter(); pend.pop(0).result(); b = time.perf_counter(); tr.append(b - a)
        while pend:
            pend.pop(0).result()
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t_all) / K * 1e3
        print(f"planes {planes} depth {depth}: {tot:.2f} ms/step; submit {sum(ts)/len(ts)*1e3:.2f} ms (max {max(ts)*1e3:.2f}), "
              f"result {sum(tr)/len(tr)*1e3:.2f} ms; chunks {CH} knobs { {k: v for k, v in os.environ.items() if k.startswith('UWCV_TRACE')} }")
