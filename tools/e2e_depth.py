"""Probe: ms/step of the streamed public call (uwcv.MeasurementStream.map, planes written, rows read
back every step) by the number of calls in flight, with pinned host inputs and with device-resident
inputs -- configs[1] on one GPU.  E2E_STEPS (default 60), E2E_DEPTHS (default "2,3,4")."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
from uwcv import _lib as _uwcv_lib
if os.environ.get("UWCV_TEST_VARIANT"):                  # e.g. "tuning": UWCV_* knobs apply
    _uwcv_lib.use_library_variant(os.environ["UWCV_TEST_VARIANT"])
import uwcv
from uwcv import synth

H = W = 2048
K = int(os.environ.get("E2E_STEPS", "60"))
DEPTHS = [int(d) for d in os.environ.get("E2E_DEPTHS", "2,3,4").split(",")]
dev = torch.device("cuda:0")
batch = synth.blob_batch(64, 1000, H, W, seed=1234)
for inst in batch:
    for k, v in list(inst.get_fields().items()):
        inst.set(k, v.pin_memory() if isinstance(v, torch.Tensor) else uwcv.Boxes(v.tensor.pin_memory()))
n = sum(len(b) for b in batch)
all_masks = torch.cat([inst.pred_masks for inst in batch]).to(dev)
dbatch, lo = [], 0
for inst in batch:
    o = uwcv.Instances(inst.image_size)
    for k, v in inst.get_fields().items():
        if k == "pred_masks":
            o.set(k, all_masks[lo:lo + len(inst)])
        else:
            o.set(k, uwcv.Boxes(v.tensor.to(dev)) if hasattr(v, "tensor") else v.to(dev))
    lo += len(inst)
    dbatch.append(o)

CASES = [(d, 4, f) for d in DEPTHS for f in ("per_chunk", "once")] 
out = []
for name, src in (("host", batch), ("device", dbatch)):
    for depth, chunks, fills in CASES:
        if name == "device" and (chunks != 4 or fills != "once"):
            continue                      # device-resident masks: one chunk whatever is asked
        st = uwcv.MeasurementStream(dev, depth=depth, fills=fills)
        for _ in st.map((src for _ in range(depth + 3)), (H, W), write_planes=True, pipeline_chunks=chunks):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = 0
        for table in st.map((src for _ in range(K)), (H, W), write_planes=True, pipeline_chunks=chunks):
            got += len(table)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / K * 1e3
        assert got == K * n
        rec = {"inputs": name, "depth": depth, "chunks": chunks, "fills": fills, "ms_per_step": round(ms, 3),
               "instances_per_s": round(n / ms * 1e3), "steps": K}
        out.append(rec)
        print(json.dumps(rec), flush=True)

# run-to-run spread of the adopted form
for rep in range(int(os.environ.get("E2E_REPS", "3"))):
    st = uwcv.MeasurementStream(dev, depth=3)
    for _ in st.map((batch for _ in range(6)), (H, W), write_planes=True):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for table in st.map((batch for _ in range(100)), (H, W), write_planes=True):
        pass
    torch.cuda.synchronize()
    print(json.dumps({"inputs": "host", "depth": 3, "fills": "default", "rep": rep, "steps": 100,
                      "ms_per_step": round((time.perf_counter() - t0) / 100 * 1e3, 3)}), flush=True)
