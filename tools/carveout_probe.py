"""Probe (tuning library): ms/step of the streamed call by the dynamic shared memory the border trace
asks for (UWCV_TRACE_SMEM) and with the carve-out hints off (UWCV_NO_CARVEOUT) -- which launch
orders leave the persistent plane fill with one CTA per SM (share_carveout, uwcv_common.cuh)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
from uwcv import _lib as _uwcv_lib
_uwcv_lib.use_library_variant("tuning")
import uwcv
from uwcv import synth

H = W = 2048
K = int(os.environ.get("E2E_STEPS", "40"))
dev = torch.device("cuda:0")
batch = synth.blob_batch(64, 1000, H, W, seed=1234)
for inst in batch:
    for k, v in list(inst.get_fields().items()):
        inst.set(k, v.pin_memory() if isinstance(v, torch.Tensor) else uwcv.Boxes(v.tensor.pin_memory()))
n = sum(len(b) for b in batch)
SMEM = [s for s in os.environ.get("PROBE_SMEM", "0,1024,16384,40960").split(",")]
for knobs in [{"UWCV_NO_CARVEOUT": "1"}] + [{"UWCV_TRACE_SMEM": s} for s in SMEM]:
    for k in ("UWCV_NO_CARVEOUT", "UWCV_TRACE_SMEM"):
        os.environ.pop(k, None)
    os.environ.update(knobs)
    for depth, fills in ((2, "once"), (2, "per_chunk"), (3, "once"), (3, "per_chunk")):
        st = uwcv.MeasurementStream(dev, depth=depth, fills=fills)
        for _ in st.map((batch for _ in range(depth + 3)), (H, W), write_planes=True):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for table in st.map((batch for _ in range(K)), (H, W), write_planes=True):
            pass
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / K * 1e3
        print(json.dumps(dict(knobs, depth=depth, fills=fills, ms_per_step=round(ms, 3), steps=K)), flush=True)
