"""Probe: CUPTI timeline (torch.profiler) of the streamed end-to-end path; prints, per step,
when the H2D copies, the kernels and the D2H copies ran.  Usage: e2e_trace.py [write_planes 0/1] [depth] [fills: per_chunk / once / default] [inputs: host / device]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch, uwcv
from torch.profiler import profile, ProfilerActivity
from uwcv import synth
H = W = 2048
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fills = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "default" else None
inputs = sys.argv[4] if len(sys.argv) > 4 else "host"
batch = synth.blob_batch(64, 1000, H, W, seed=1234)
for inst in batch:
    for k, v in list(inst.get_fields().items()):
        inst.set(k, v.pin_memory() if isinstance(v, torch.Tensor) else uwcv.Boxes(v.tensor.pin_memory()))
if inputs == "device":
    all_masks = torch.cat([inst.pred_masks for inst in batch]).cuda()
    dbatch, lo = [], 0
    for inst in batch:
        o = uwcv.Instances(inst.image_size)
        for k, v in inst.get_fields().items():
            if k == "pred_masks":
                o.set(k, all_masks[lo:lo + len(inst)])
            else:
                o.set(k, uwcv.Boxes(v.tensor.cuda()) if hasattr(v, "tensor") else v.cuda())
        lo += len(inst)
        dbatch.append(o)
    batch = dbatch
st = uwcv.MeasurementStream(depth=depth, fills=fills)
for _ in st.map((batch for _ in range(depth + 4)), (H, W), write_planes=bool(planes)): pass
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    for _ in st.map((batch for _ in range(8)), (H, W), write_planes=bool(planes)): pass
    torch.cuda.synchronize()
    print("profiled ms/step", (time.perf_counter() - t0) / 8 * 1e3)
out = os.path.join(ROOT, "gpurun_out", f"trace_p{planes}_d{depth}_{fills}_{inputs}.json")
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
gpu.sort(key=lambda e: e["ts"])
base = gpu[0]["ts"]
rows = []
for e in gpu:
    name = e["name"]
    kind = "H2D" if "HtoD" in name else "D2H" if "DtoH" in name else name.split("(")[0][-40:]
    rows.append((e["ts"] - base, e["dur"], kind, e.get("args", {}).get("stream")))
# merge runs of the same kind on the same stream
merged = []
for ts, dur, kind, s in rows:
    if merged and merged[-1][2] == kind and merged[-1][3] == s and ts - (merged[-1][0] + merged[-1][1]) < 200:
        m = merged[-1]
        merged[-1] = (m[0], ts + dur - m[0], kind, s, m[4] + 1, m[5] + dur)
    else:
        merged.append((ts, dur, kind, s, 1, dur))
with open(os.path.join(ROOT, "gpurun_out", f"trace_p{planes}_d{depth}_{fills}_{inputs}.txt"), "w") as f:
    for m in merged:
        f.write(f"{m[0]/1e3:9.3f} ms  +{m[1]/1e3:7.3f} ms  busy {m[5]/1e3:7.3f}  x{m[4]:<3d} stream {m[3]}  {m[2]}\n")
os.remove(out)
