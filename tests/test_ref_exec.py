"""CPU: the oracle restatements against outputs of the reference's OWN functions.

``tests/golden/ref_exec_*.npz`` hold what ``/root/reference/nn_inference.py``'s
``GetMask_Contours`` / ``GetCounts`` / driver loop / ``postprocess_masks`` / ``rle_encoding`` /
export loop returned when executed (AST nodes, unmodified; oracle/ref_exec.py) on the seeded inputs
of ``tests/ref_fixtures.py``.  These tests pin ``oracle/measure.py``, ``oracle/cleanup.py`` and the
host report layer to them bit for bit -- an edit to ``oracle.measure.contour_descriptors`` fails
here -- and, where ``/root/reference`` is present, re-execute the reference live to prove that the
committed goldens are what it produces today."""
import json
import os

import numpy as np
import pytest
import torch

import ref_fixtures as FX
import uwcv
from oracle import cleanup as OC, d2, measure as M, pipeline as P, ref_exec as R

KEYWORDS = ["Scale", "WThick", "PThroat", "Pore"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "ref_exec_manifest.json")) as f:
        man = json.load(f)
    return (man, np.load(os.path.join(golden_dir, "ref_exec_measure.npz")),
            np.load(os.path.join(golden_dir, "ref_exec_cleanup.npz")))


def _fixture(tag, golden_dir):
    return {"union_dense": FX.union_dense, "blobs_rescaled": FX.blobs_rescaled,
            "c1_maskrcnn": lambda: FX.c1_maskrcnn(golden_dir)}[tag]()


def test_manifest_cites_the_reference_lines(gold):
    man = gold[0]
    src = man["source"]["nn_inference.py"]
    assert src["GetMask_Contours"] == [371, 459] and src["midpoint"] == [339, 340]
    assert src["GetCounts"] == [355, 366] and src["postprocess_masks"] == [265, 306]
    assert src["rle_encoding"] == [253, 263] and src["rle_decode"] == [237, 251]
    assert src["class_driver"] == [487, 570] and src["export_loop"][0] == 319
    assert man["source"]["backup_main.py"]["GetMask_Contours"] == [429, 497]
    # the entry types the moving average depends on (float32 lists)
    d = man["measure"]["blobs_rescaled"]["dtypes"]
    got = [d[n] for n in R.CSV_ORDER]
    want = {np.float32: "float32", np.float64: "float64", float: "float"}
    assert got == [want[t] for t in M.CSV_ENTRY_TYPES]
    assert tuple({"float32": "f32", "float64": "f64", "float": "py"}[g] for g in got) == \
        uwcv.schema.CSV_KINDS


@pytest.mark.parametrize("tag", ["union_dense", "blobs_rescaled", "c1_maskrcnn"])
def test_oracle_get_mask_contours_equals_the_reference(gold, golden_dir, tag):
    man, meas, _ = gold
    batch, (H, W) = _fixture(tag, golden_dir)
    info = man["measure"][tag]
    assert FX.digest(batch) == info["digest"], "fixture inputs changed: regenerate the goldens"
    total = 0
    for k, inst in enumerate(batch):
        res = d2.detector_postprocess(P.to_oracle_instances(inst), H, W, 0.5)
        for key, classes in [(f"cls{c}", [c]) for c in range(4)] + [("all", [0, 1, 2, 3])]:
            want = meas[f"{tag}/img{k}.tif/{key}"]
            err = info["errors"].get(f"img{k}.tif/{key}", "")
            try:
                got = M.get_mask_contours((H, W, 3), res.pred_classes.numpy(),
                                          res.pred_masks.numpy(), classes)
                got = np.zeros((0, 9)) if got is None else got
                assert not err
            except ValueError:
                assert err == "ValueError"
                continue
            assert got.shape == want.shape and np.array_equal(got, want), (tag, k, key)
            total += len(want)
        assert np.array_equal(meas[f"{tag}/img{k}.tif/backup_main"], meas[f"{tag}/img{k}.tif/all"])
        c = M.get_counts(res.pred_classes.numpy())
        ref = info["counts"][f"img{k}.tif"]
        assert [c["SCount"], c["WTCount"], c["PTCount"], c["PCount"]] == \
            [ref["SList"], ref["WTList"], ref["PTList"], ref["PList"]]
    assert total > 20


def test_three_ellipses_kat(gold):
    """SURVEY.md 8(c)(iv), now produced by the reference's own GetMask_Contours."""
    rows = gold[1]["three_ellipses/rows"]
    survey = np.array([
        (59.116913, 2.077556, 0.481335, 0.602998, 0.776530, 28.455027, 59.116913, 33.396974, 124.568542),
        (92.756706, 1.0, 1.0, 0.754144, 0.868415, 92.756706, 92.756706, 79.083201, 263.764500),
        (70.832909, 3.006643, 0.332597, 0.483796, 0.695555, 23.558804, 70.832909, 33.358827, 138.911687)])
    assert np.allclose(rows, survey, rtol=2e-6, atol=0)
    masks, classes = FX.three_ellipses()
    assert np.array_equal(M.get_mask_contours((200, 200, 3), classes, masks, [0]), rows)


def test_driver_loop_report_layer(gold, golden_dir):
    """nn_inference.py:487-570 for each keyword over the three fixture images: the rows its
    lists collect == the oracle's per-image rows concatenated, its ShapeDescriptor.csv == the
    oracle's and the host report layer's text, its loop ends in the documented IndexError."""
    man, meas, _ = gold
    drv = man["measure"]["blobs_rescaled"]["driver"]
    for c, kw in enumerate(KEYWORDS):
        rows = meas[f"blobs_rescaled/driver/{kw}"]
        per_image = [meas[f"blobs_rescaled/img{k}.tif/cls{c}"] for k in range(3)]
        assert np.array_equal(rows, np.concatenate(per_image))
        assert drv[kw]["ended"].startswith("IndexError") and drv[kw]["count"] == 3
        text = drv[kw]["shape_csv"]
        assert M.shape_descriptor_text(rows) == text
        p = os.path.join(golden_dir, "..", "_shape_tmp.csv")
        try:
            uwcv.write_shape_descriptor_csv(p, rows)
            with open(p) as f:
                assert f.read() == text
        finally:
            if os.path.exists(p):
                os.remove(p)
        sm, _ = uwcv.report_class(rows)
        cells = [[float(v) for v in line.split(",")] for line in text.split()]
        # float32 cells print as their shortest float32 text: compare at float32 where typed so
        for j, kind in enumerate(uwcv.schema.CSV_KINDS):
            col = np.array([r[j] for r in cells])
            mine = sm[:, j]
            if kind == "f32":
                assert np.array_equal(col.astype(np.float32), mine.astype(np.float32)), (kw, j)
            else:
                assert np.array_equal(col, mine), (kw, j)
        # counts: one GetCounts per image and keyword (:492), totals accumulate (:541-548)
        imgs = man["measure"]["blobs_rescaled"]["counts"]
        assert drv[kw]["counts"]["SList"] == [imgs[f"img{k}.tif"]["SList"] for k in range(3)]
        assert drv[kw]["totals"]["tPT"] == sum(drv[kw]["counts"]["PTList"])


def test_oracle_postprocess_masks_and_rle_equal_the_reference(gold):
    man, _, cl = gold
    info = man["cleanup"]["postprocess_masks"]
    n_list = 0
    for case in FX.bool_mask_cases():
        H, W = case["masks"].shape[1:]
        want = info[case["name"]]
        got = OC.postprocess_masks(case["masks"].copy(), case["scores"].copy(), (H, W))
        kind = "None" if got is None else ("empty" if len(got) == 0 else "list")
        assert kind == want["kind"], case["name"]
        if kind == "list":
            ref = np.unpackbits(cl[f"pp/{case['name']}"], axis=-1)[..., :W]
            assert len(got) == want["n"] == ref.shape[0]
            assert np.array_equal(np.stack(got), ref), case["name"]
            n_list += 1
    assert n_list >= 9
    for k, x in enumerate(FX.rle_cases()):
        ref = cl[f"rle/{k}"].tolist()
        assert OC.rle_encoding(x) == ref and OC.rle_encoding_literal(x) == ref
        if ref:
            dec = OC.rle_decode(' '.join(map(str, ref)), (x.shape[1], x.shape[0])).T
            assert np.array_equal(dec, x)
    assert all(man["cleanup"]["rle_round_trip"])


def test_oracle_export_rows_equal_the_reference_csv(gold):
    man, _, cl = gold
    batch, names, (H, W) = FX.export_batch()
    exp = man["cleanup"]["export"]
    assert FX.digest(batch) == exp["digest"], "fixture inputs changed: regenerate the goldens"
    masks_l, scores_l = [], []
    for b in batch:
        res = d2.detector_postprocess(P.to_oracle_instances(b), H, W, 0.5)
        masks_l.append(res.pred_masks.numpy())
        scores_l.append(res.scores.numpy())
    ids, enc = OC.export_rows(names, masks_l, scores_l, (H, W))
    text = bytes(cl["export/csv"]).decode()
    lines = text.split("\n")
    assert lines[0] == "ImageId,EncodedPixels" and lines[-1] == ""
    assert ids == exp["image_ids"] and len(lines) - 2 == len(ids)
    for k, line in enumerate(lines[1:-1]):
        assert line == f"{ids[k]},{enc[k]}", k


@pytest.mark.skipif(not R.available(), reason="/root/reference is only in the build container")
def test_goldens_are_what_the_reference_produces_today(gold, golden_dir):
    """Live re-execution of the reference's nodes on two fixtures == the committed goldens."""
    man, meas, cl = gold
    assert R.node_lines(R.NN_INFERENCE) == {k: tuple(v) for k, v in man["source"]["nn_inference.py"].items()}
    batch, (H, W) = FX.blobs_rescaled()
    images = {f"img{k}.tif": (np.zeros((H, W, 3), np.uint8),
                              d2.detector_postprocess(P.to_oracle_instances(b), H, W, 0.5))
              for k, b in enumerate(batch)}
    r = R.ReferenceRunner(images)
    for c in range(4):
        assert np.array_equal(r.get_mask_contours("img1.tif", [c]), meas[f"blobs_rescaled/img1.tif/cls{c}"])
    out = r.run_class_driver("Pore")
    assert out["shape_csv"] == man["measure"]["blobs_rescaled"]["driver"]["Pore"]["shape_csv"]
    case = FX.bool_mask_cases()[0]
    H, W = case["masks"].shape[1:]
    got = r.call("postprocess_masks", case["masks"].copy(), case["scores"].copy(), np.zeros((H, W, 3), np.uint8))
    assert np.array_equal(np.stack(got), np.unpackbits(cl[f"pp/{case['name']}"], axis=-1)[..., :W])
