"""Work distribution of the border-trace kernel on configs[1] (tuning library): per instance the
scan steps, border steps, contours and the loop iteration at which its lane finished."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from uwcv import _lib  # noqa: E402
_lib.use_library_variant("tuning")
from uwcv import api, synth  # noqa: E402

images = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = W = 2048
dev = torch.device("cuda", 0)
batch = synth.blob_batch(images, 1000, H, W, seed=1234)
eng = api.Engine.get(dev)
boxes = torch.cat([api.scale_clip_boxes(b.pred_boxes.tensor, (H, W), (H, W))[0] for b in batch])
n = int(boxes.shape[0])
d_boxes = boxes.to(dev)
d_masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous().to(dev)
stats = torch.zeros((n, 4), dtype=torch.int32, device=dev)
L = _lib.lib()
L.uwcv_tuning_set_trace_stats.argtypes = [C.c_void_p]
assert L.uwcv_tuning_set_trace_stats(stats.data_ptr()) == 0
ri, rf, st = eng.run(d_masks, d_boxes, H, W)
torch.cuda.synchronize()
s = stats.cpu().numpy().astype(np.int64)
tot = s[:, 0] + s[:, 1]
q = [50, 90, 99, 99.9, 100]
out = {"n": n, "scan_sum": int(s[:, 0].sum()), "trace_sum": int(s[:, 1].sum()),
       "scan_pct": np.percentile(s[:, 0], q).tolist(), "trace_pct": np.percentile(s[:, 1], q).tolist(),
       "total_pct": np.percentile(tot, q).tolist(), "ncont_pct": np.percentile(s[:, 2], q).tolist(),
       "end_iter_pct": np.percentile(s[:, 3], q).tolist(),
       "multi_contour_frac": float((s[:, 2] > 1).mean())}
# the 10 longest lanes
top = np.argsort(-tot)[:10]
hi = ri.cpu().numpy()
out["longest"] = [dict(inst=int(i), scan=int(s[i, 0]), trace=int(s[i, 1]), ncont=int(s[i, 2]),
                       area=int(hi[i, 5]), bbox=hi[i, 6:10].tolist()) for i in top]
# warp view: 22 consecutive instances per warp, the warp lasts as long as its longest lane
lanes = 22
w = tot[: n // lanes * lanes].reshape(-1, lanes)
out["warp_max_mean"] = float(w.max(1).mean())
out["warp_sum_over_max"] = float((w.sum(1) / np.maximum(w.max(1), 1)).mean())
words_each = api.tile_words_each(boxes, H, W).numpy()
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "trace_stats_arrays.npz"), stats=s, words=words_each,
                    area=hi[:, 5], bbox=hi[:, 6:10], boxes=boxes.numpy())
print(json.dumps(out))
