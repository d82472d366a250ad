"""Randomised parity sweep (GPU box): measure_instances and export_rle against the CPU oracle over
random image sizes, box distributions and mask sources.  Prints one JSON summary.
    python tools/parity_sweep.py [n_cases] [seed0]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from uwcv import _lib as _uwcv_lib
if os.environ.get('UWCV_TEST_VARIANT'):
    _uwcv_lib.use_library_variant(os.environ['UWCV_TEST_VARIANT'])
import uwcv
from uwcv import synth, schema
from oracle import d2, pipeline as P, cleanup as OC

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
IC, FC = schema.ICOL, schema.FCOL
tot = dict(cases=0, instances=0, int_cells=0, int_mismatch=0, float_cells=0, float_exact=0,
           float_over_1e5=0, rle_rows=0, rle_mismatch=0, plane_crc_mismatch=0)
t0 = time.time()
for case in range(n_cases):
    g = torch.Generator().manual_seed(seed0 + case)
    H = int(torch.randint(40, 700, (1,), generator=g)); W = int(torch.randint(40, 700, (1,), generator=g))
    n = int(torch.randint(1, 90, (1,), generator=g))
    kind = case % 4
    inst = synth.blob_instances(case, n, H, W, seed=seed0 + case)
    n = len(inst)
    if kind == 1:        # saturated / binary probabilities: threshold ties, speckle
        p = inst.pred_masks
        inst.set("pred_masks", torch.where(torch.rand(p.shape, generator=g) < 0.5, (p > 0.5).float(), p))
    elif kind == 2:      # sub-pixel and huge boxes, boxes crossing the border
        c = torch.rand((n, 2), generator=g) * torch.tensor([W, H]) * 1.2 - 0.1 * torch.tensor([W, H])
        wh = torch.exp(torch.rand((n, 2), generator=g) * 8 - 2.5)
        inst.set("pred_boxes", uwcv.Boxes(torch.cat((c - wh / 2, c + wh / 2), 1).float()))
    elif kind == 3:      # network-input size differs from the output size (anisotropic rescale)
        inst = uwcv.Instances((int(H * 0.77) + 3, int(W * 1.31) + 5), **{k: v for k, v in inst.get_fields().items()})
    order = torch.argsort(inst.scores, descending=True)
    inst2 = uwcv.Instances(inst.image_size)
    for k, v in inst.get_fields().items():
        inst2.set(k, uwcv.Boxes(v.tensor[order]) if hasattr(v, "tensor") else v[order])
    inst = inst2
    table, planes = uwcv.measure_instances(inst, (H, W), return_planes=True)
    rows_only = uwcv.measure_instances(inst, (H, W))       # no planes: one instance per warp
    tot["rows_only_mismatch"] = tot.get("rows_only_mismatch", 0) + int(
        not (np.array_equal(rows_only.ints, table.ints) and
             np.array_equal(rows_only.floats, table.floats, equal_nan=True)))
    ri, rf = P.oracle_table([inst], (H, W))
    assert table.ints.shape == ri.shape, (case, table.ints.shape, ri.shape)
    tot["cases"] += 1; tot["instances"] += len(ri)
    tot["int_cells"] += ri.size; tot["int_mismatch"] += int((table.ints != ri).sum())
    a, b = table.floats, rf
    ok = ~np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), case
    tot["float_cells"] += int(ok.sum()); tot["float_exact"] += int((a[ok] == b[ok]).sum())
    m00 = np.maximum(ri[:, IC["area_px"]].astype(np.float64), 1.0)[:, None] * np.ones_like(b)
    scale = np.abs(b)
    for name in ("mu30", "mu21", "mu12", "mu03"):
        scale[:, FC[name]] = np.maximum(scale[:, FC[name]], m00[:, FC[name]] ** 2.5 * 1e-6)
    scale[:, FC["mu11"]] = np.maximum(scale[:, FC["mu11"]], 1e-6 * m00[:, FC["mu11"]] ** 2)
    scale[:, FC["ell_theta"]] = np.maximum(scale[:, FC["ell_theta"]], 1e-6)
    over = ok & (np.abs(a - b) > 1e-5 * scale + 1e-300)
    tot["float_over_1e5"] += int(over.sum())
    for r_, c_ in zip(*np.nonzero(over)):          # which cells: (case, row, column, device, oracle)
        if len(tot.setdefault("float_over_cells", [])) < 40:
            tot["float_over_cells"].append([case, int(r_), schema.FLOAT_COLUMNS[c_] if hasattr(schema, "FLOAT_COLUMNS")
                                            else int(c_), float(a[r_, c_]), float(b[r_, c_])])
    # planes against the oracle's pasted masks
    res = d2.detector_postprocess(P.to_oracle_instances(inst), H, W, 0.5)
    om = res.pred_masks.numpy()
    pl = planes.cpu().numpy().view(np.uint32) if planes is not None else np.zeros((0, H, 4), np.uint32)
    assert len(pl) == len(om), (case, len(pl), len(om))
    for i in range(len(om)):
        if not np.array_equal(pl[i], P.pack_bits(om[i], pl.shape[2])):
            tot["plane_crc_mismatch"] += 1
    ids, enc = OC.export_rows(["x.tif"], [om], [res.scores.numpy()], (H, W))
    got = uwcv.export_rle(inst, (H, W), ["x.tif"])
    tot["rle_rows"] += len(enc)
    tot["rle_mismatch"] += int(got.image_id != ids) + sum(1 for x, y in zip(got.encoded_pixels, enc) if x != y) \
        + abs(len(got.encoded_pixels) - len(enc))
tot["seconds"] = round(time.time() - t0, 1)
tot["seed0"] = seed0
print(json.dumps(tot))
