"""Freeze outputs of the reference's OWN functions (build container only).

Runs ``/root/reference/nn_inference.py``'s ``GetMask_Contours`` (:371-459, with ``midpoint``
:339-340), ``GetCounts`` (:355-366), the per-keyword driver loop (:487-570: moving average,
ShapeDescriptor.csv), ``postprocess_masks`` (:265-306), ``rle_encoding`` (:253-263), ``rle_decode``
(:237-251), the export loop (:313-336) and ``backup_main.py``'s ``GetMask_Contours`` (:429-497, no
class filter) through ``oracle/ref_exec.py`` -- AST nodes taken from the files where they lie,
compiled unmodified, run with stubs for what the image lacks (see that module's header) -- on the
seeded inputs of ``tests/ref_fixtures.py`` and writes

  tests/golden/ref_exec_measure.npz    K x 9 rows per (fixture, image, class set)
  tests/golden/ref_exec_cleanup.npz    cleaned masks (bit-packed), RLE lists, export strings
  tests/golden/ref_exec_manifest.json  source line ranges of the executed nodes, library
                                       versions, list entry types, counts, driver-loop results,
                                       input digests

``/root/reference`` does not travel to the GPU box; these files do.
Usage:  python tests/golden/make_ref_golden.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "uw-com-vision_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cv2  # noqa: E402
import scipy  # noqa: E402
import ref_fixtures as FX  # noqa: E402
from oracle import d2, pipeline as P, ref_exec as R  # noqa: E402

KEYWORDS = ["Scale", "WThick", "PThroat", "Pore"]


def predictor_outputs(batch, out_size):
    """What the stub ``predictor(im)`` hands the reference: post-processed Instances."""
    H, W = out_size
    return {f"img{k}.tif": (np.zeros((H, W, 3), np.uint8),
                            d2.detector_postprocess(P.to_oracle_instances(inst), H, W, 0.5))
            for k, inst in enumerate(batch)}


def rows_or_error(fn):
    try:
        return fn(), ""
    except Exception as e:                      # imutils.sort_contours on an empty contour list
        return np.zeros((0, 9)), f"{type(e).__name__}"


def measure_fixture(tag, batch, out_size, arrays, manifest, driver=False):
    images = predictor_outputs(batch, out_size)
    r = R.ReferenceRunner(images)
    info = dict(out_size=list(out_size), digest=FX.digest(batch), errors={}, counts={}, dtypes={})
    for name in images:
        for cls in range(4):
            rows, err = rows_or_error(lambda: r.get_mask_contours(name, [cls]))
            arrays[f"{tag}/{name}/cls{cls}"] = rows
            if err:
                info["errors"][f"{name}/cls{cls}"] = err
            elif len(rows) and not info["dtypes"]:
                info["dtypes"] = r.list_dtypes()
        rows, err = rows_or_error(lambda: r.get_mask_contours(name, [0, 1, 2, 3]))
        arrays[f"{tag}/{name}/all"] = rows
        if err:
            info["errors"][f"{name}/all"] = err
        info["counts"][name] = r.get_counts(name)
    # second witness: backup_main.py's GetMask_Contours() has no class filter
    rb = R.ReferenceRunner(images, path=R.BACKUP_MAIN)
    for name in images:
        rows, err = rows_or_error(lambda: rb.get_mask_contours(name, None))
        arrays[f"{tag}/{name}/backup_main"] = rows
        if not err:
            assert np.array_equal(rows, arrays[f"{tag}/{name}/all"]), "the two scripts disagree"
    if driver:
        info["driver"] = {}
        for kw in KEYWORDS:
            out = r.run_class_driver(kw)
            arrays[f"{tag}/driver/{kw}"] = out["rows"]
            info["driver"][kw] = {k: out[k] for k in ("ended", "counts", "totals", "shape_csv",
                                                      "count", "dtypes")}
    manifest["measure"][tag] = info
    n = sum(len(v) for k, v in arrays.items() if k.startswith(tag + "/") and "/cls" in k)
    print(f"{tag}: {n} per-class rows, errors {info['errors']}")


def main():
    assert R.available(), "needs /root/reference (build container)"
    d2.assert_cpu_capability()
    manifest = dict(
        source={"nn_inference.py": R.node_lines(R.NN_INFERENCE),
                "backup_main.py": R.node_lines(R.BACKUP_MAIN)},
        versions=dict(numpy=np.__version__, cv2=cv2.__version__, scipy=scipy.__version__,
                      torch=torch.__version__),
        substitutions=["predictor -> oracle/d2.detector_postprocess (Detectron2 absent)",
                       "imutils -> oracle/imutils_port (imutils absent)",
                       "skimage erosion/dilation/label -> scipy.ndimage grey_erosion/grey_dilation"
                       "/label (scikit-image absent)"],
        measure={}, cleanup={})
    arrays = {}
    batch, size = FX.union_dense()
    measure_fixture("union_dense", batch, size, arrays, manifest)
    batch, size = FX.blobs_rescaled()
    measure_fixture("blobs_rescaled", batch, size, arrays, manifest, driver=True)
    batch, size = FX.c1_maskrcnn(HERE)
    measure_fixture("c1_maskrcnn", batch, size, arrays, manifest)
    # masks given directly
    masks, classes = FX.three_ellipses()
    inst = d2.Instances((200, 200))
    inst.pred_masks = torch.from_numpy(masks)
    inst.pred_classes = torch.from_numpy(classes)
    r = R.ReferenceRunner({"kat.tif": (np.zeros((200, 200, 3), np.uint8), inst)})
    arrays["three_ellipses/rows"] = r.get_mask_contours("kat.tif", [0])
    np.savez_compressed(os.path.join(HERE, "ref_exec_measure.npz"), **arrays)

    # ---- clean-up + RLE ---------------------------------------------------------------
    carr = {}
    r = R.ReferenceRunner({})
    info = {}
    for case in FX.bool_mask_cases():
        H, W = case["masks"].shape[1:]
        out = r.call("postprocess_masks", case["masks"].copy(), case["scores"].copy(),
                     np.zeros((H, W, 3), np.uint8))
        kind = "None" if out is None else ("empty" if len(out) == 0 else "list")
        info[case["name"]] = dict(kind=kind, n=0 if out is None else len(out), shape=[H, W])
        if kind == "list":
            carr[f"pp/{case['name']}"] = np.packbits(np.stack(out).astype(bool), axis=-1)
    manifest["cleanup"]["postprocess_masks"] = info
    rl = []
    for k, x in enumerate(FX.rle_cases()):
        enc = r.call("rle_encoding", x)
        carr[f"rle/{k}"] = np.asarray(enc, dtype=np.int64)
        dec = r.call("rle_decode", ' '.join(map(str, enc)), (x.shape[1], x.shape[0])).T
        rl.append(bool(np.array_equal(dec, x)))
    manifest["cleanup"]["rle_round_trip"] = rl
    batch, names, size = FX.export_batch()
    H, W = size
    images = {n: (np.zeros((H, W, 3), np.uint8),
                  d2.detector_postprocess(P.to_oracle_instances(b), H, W, 0.5))
              for n, b in zip(names, batch)}
    r = R.ReferenceRunner(images)
    ids, enc, text = r.run_export_loop()
    carr["export/csv"] = np.frombuffer(text.encode(), dtype=np.uint8)
    manifest["cleanup"]["export"] = dict(size=[H, W], names=names, rows=len(ids),
                                         digest=FX.digest(batch), image_ids=ids)
    assert text.count("\n") == len(ids) + 1
    np.savez_compressed(os.path.join(HERE, "ref_exec_cleanup.npz"), **carr)
    with open(os.path.join(HERE, "ref_exec_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("export:", len(ids), "rows;", "postprocess cases:", {k: v["kind"] for k, v in info.items()})


if __name__ == "__main__":
    main()
