"""Probe: where the end-to-end step goes (host submit / result times, stream vs synchronous)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch, uwcv
from uwcv import synth
H = W = 2048
batch = synth.blob_batch(64, 1000, H, W, seed=1234)
for inst in batch:
    for k, v in list(inst.get_fields().items()):
        inst.set(k, v.pin_memory() if isinstance(v, torch.Tensor) else uwcv.Boxes(v.tensor.pin_memory()))
for _ in range(3):
    uwcv.measure_instances(batch, (H, W), write_planes=True)
torch.cuda.synchronize()
K = 10
t0 = time.perf_counter()
for _ in range(K):
    uwcv.measure_instances(batch, (H, W), write_planes=True)
print("sync ms/call", (time.perf_counter() - t0) / K * 1e3)
for depth in (1, 2, 3):
    for chunks in (2, 4, 8):
        st = uwcv.MeasurementStream(depth=depth)
        for _ in st.map((batch for _ in range(3)), (H, W), write_planes=True, pipeline_chunks=chunks):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in st.map((batch for _ in range(K)), (H, W), write_planes=True, pipeline_chunks=chunks):
            pass
        torch.cuda.synchronize()
        print(f"stream depth {depth} chunks {chunks}: ms/step", (time.perf_counter() - t0) / K * 1e3)
# host cost of submit and result
st = uwcv.MeasurementStream(depth=2)
ts, tr = [], []
pend = None
for i in range(K):
    a = time.perf_counter(); p = st.submit(batch, (H, W), write_planes=True); b = time.perf_counter()
    ts.append(b - a)
    if pend is not None:
        a = time.perf_counter(); pend.result(); b = time.perf_counter(); tr.append(b - a)
    pend = p
pend.result()
print("submit host ms", [round(x * 1e3, 2) for x in ts])
print("result host ms", [round(x * 1e3, 2) for x in tr])
# cropped (no planes) for comparison
st = uwcv.MeasurementStream(depth=2)
for _ in st.map((batch for _ in range(3)), (H, W)): pass
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in st.map((batch for _ in range(K)), (H, W)): pass
torch.cuda.synchronize()
print("stream depth 2, no planes: ms/step", (time.perf_counter() - t0) / K * 1e3)
m = torch.cat([b.pred_masks for b in batch]).pin_memory()
d = torch.empty_like(m, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(m, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"H2D {m.numel()*4/1e6:.0f} MB in {dt*1e3:.2f} ms = {m.numel()*4/dt/1e9:.1f} GB/s")
