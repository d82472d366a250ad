"""Probe: host-side profile of uwcv.measure_instances on the bench workload."""
import cProfile, pstats, os, sys, time
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
t0 = time.perf_counter()
for _ in range(5):
    uwcv.measure_instances(batch, (H, W), write_planes=True)
print("ms/call", (time.perf_counter() - t0) / 5 * 1e3)
# raw H2D bandwidth from pinned
m = torch.cat([b.pred_masks for b in batch]).pin_memory()
d = torch.empty_like(m, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(m, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"H2D {m.numel()*4/1e6:.0f} MB in {dt*1e3:.2f} ms = {m.numel()*4/dt/1e9:.1f} GB/s")
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    uwcv.measure_instances(batch, (H, W), write_planes=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
