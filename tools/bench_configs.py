"""Timings of the other BASELINE.json configs (parity-test shapes, not bench lines):
config 1 (1024x1024, 200 raw Mask R-CNN instances), config 2 in cropped mode (rows only),
config 4 (4096x4096, ~20 k candidates -> NMS -> ~5 k instances).  Prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import numpy as np, torch
import uwcv
from uwcv import api, synth

dev = torch.device("cuda", 0)
eng = api.Engine.get(dev)
out = {}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps


def device_inputs(batch, H, W):
    boxes = torch.cat([api.scale_clip_boxes(b.pred_boxes.tensor, b.image_size, (H, W))[0] for b in batch])
    masks = torch.cat([b.pred_masks[:, 0] for b in batch]).contiguous()
    return boxes.to(dev), masks.to(dev), api.tile_words(boxes, H, W)

# ---- config 1 -----------------------------------------------------------------------------
g = np.load(os.path.join(ROOT, "tests", "golden", "c1_maskrcnn.npz"))
inst = uwcv.Instances((1024, 1024), pred_boxes=uwcv.Boxes(torch.from_numpy(g["boxes"])),
                      scores=torch.from_numpy(g["scores"]), pred_classes=torch.from_numpy(g["classes"]),
                      pred_masks=torch.from_numpy(g["masks"]))
b, m, words = device_inputs([inst], 1024, 1024)
n = len(b)
planes = eng.alloc_planes(n, 1024, 1024)
ri = torch.empty((n, 20), dtype=torch.int64, device=dev); rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
ms = timed(lambda: eng.run(m, b, 1024, 1024, planes=planes, n_tile_words=words, rows_i=ri, rows_f=rf))
ms_c = timed(lambda: eng.run(m, b, 1024, 1024, n_tile_words=words, rows_i=ri, rows_f=rf))
t0 = time.perf_counter(); uwcv.measure_instances(inst, (1024, 1024), write_planes=True); e2e = (time.perf_counter() - t0) * 1e3
out["config1_1024_200_raw_maskrcnn"] = {"instances": n, "device_ms_full_frame": ms, "device_ms_cropped": ms_c,
                                       "e2e_ms_single_call": e2e, "instances_per_s_full_frame": n / ms * 1e3}
# ---- config 2, cropped contract -------------------------------------------------------------
batch = synth.blob_batch(64, 1000, 2048, 2048, seed=1234)
b, m, words = device_inputs(batch, 2048, 2048)
n = len(b)
ri = torch.empty((n, 20), dtype=torch.int64, device=dev); rf = torch.empty((n, 30), dtype=torch.float64, device=dev)
ms_c = timed(lambda: eng.run(m, b, 2048, 2048, n_tile_words=words, rows_i=ri, rows_f=rf))
out["config2_cropped_rows_only"] = {"instances": n, "device_ms": ms_c, "instances_per_s": n / ms_c * 1e3,
                                    "algorithmic_bytes_per_instance": 3168 + 400 + words * 4 / n,
                                    "achieved_gbs": n * (3168 + 400 + words * 4 / n) / (ms_c * 1e-3) / 1e9}
# ---- config 4 ---------------------------------------------------------------------------------
H = W = 4096
cb, cs, cc = synth.clustered_candidates(5000, H, W, seed=99)
dcb, dcs, dcc = cb.to(dev), cs.to(dev), cc.to(dev)
ms_nms = timed(lambda: eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000), reps=5)
keep, cnt = eng.nms(dcb, dcs, dcc, [0, len(cb)], 0.05, 0.5, 6000)
k = int(cnt[0]); keep = keep[:k].cpu()
gg = torch.Generator().manual_seed(8)
inst = uwcv.Instances((H, W), pred_boxes=uwcv.Boxes(cb[keep]), scores=cs[keep], pred_classes=cc[keep],
                      pred_masks=synth.blob_probs(k, gg)[:, None])
b, m, words = device_inputs([inst], H, W)
planes = eng.alloc_planes(k, H, W)
ri = torch.empty((k, 20), dtype=torch.int64, device=dev); rf = torch.empty((k, 30), dtype=torch.float64, device=dev)
ms = timed(lambda: eng.run(m, b, H, W, planes=planes, n_tile_words=words, rows_i=ri, rows_f=rf), reps=5)
ms_c = timed(lambda: eng.run(m, b, H, W, n_tile_words=words, rows_i=ri, rows_f=rf))
bpi = 3168 + 400 + H * W // 8
out["config4_4096_dense"] = {"candidates": len(cb), "instances": k, "nms_ms": ms_nms,
                             "device_ms_full_frame": ms, "device_ms_cropped": ms_c,
                             "full_frame_gbs": k * bpi / (ms * 1e-3) / 1e9,
                             "instances_per_s_full_frame": k / ms * 1e3}
# ---- union mode on one config-2 image -----------------------------------------------------------
def wall(fn, reps=5, warm=2):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return r, sorted(ts)[len(ts) // 2]


ut, ms_u = wall(lambda: uwcv.measure_union(batch[:8], (2048, 2048), classes_of_interest=[3]))
out["union_mode_8_images_class3"] = {"rows": len(ut), "wall_ms": ms_u}
# ---- f2: mask clean-up + RLE export on 8 config-2 images ------------------------------------------
ex, ms_x = wall(lambda: uwcv.export_rle(batch[:8], (2048, 2048)))
out["rle_export_8_images"] = {"rows": len(ex), "wall_ms": ms_x,
                              "emptied_multi_piece": int(ex.multi_piece.sum())}
print(json.dumps(out, indent=1))
