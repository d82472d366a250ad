"""GPU: the CUDA path (through the C ABI) against outputs of the reference's OWN functions.

The goldens under tests/golden/ref_exec_* were produced by executing the reference's
``GetMask_Contours`` / ``GetCounts`` / driver loop / ``postprocess_masks`` / ``rle_encoding`` / export
loop (nn_inference.py:237-306, :313-336, :339-366, :371-459, :487-570; see
tests/golden/make_ref_golden.py) on the inputs of tests/ref_fixtures.py.  Nothing of the oracle's
measurement / clean-up code is on the checking side here: golden file in, CUDA rows out.

Bar: row counts, instance counts, cleaned masks and EncodedPixels strings identical; the nine float
columns within 1e-6 relative (north_star allows 1e-5)."""
import json
import os

import numpy as np
import pytest
import torch

import ref_fixtures as FX
import uwcv

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "ref_exec_manifest.json")) as f:
        man = json.load(f)
    return (man, np.load(os.path.join(golden_dir, "ref_exec_measure.npz")),
            np.load(os.path.join(golden_dir, "ref_exec_cleanup.npz")))


def _close(mine, want, what):
    assert mine.shape == want.shape, (what, mine.shape, want.shape)
    if len(want):
        rel = np.abs(mine - want) / np.maximum(np.abs(want), 1e-30)
        assert rel.max() <= RTOL, (what, rel.max(axis=0))
    return int(np.array_equal(mine, want)), len(want)


@pytest.mark.parametrize("tag", ["union_dense", "blobs_rescaled", "c1_maskrcnn"])
def test_union_rows_equal_the_reference_get_mask_contours(gold, golden_dir, tag):
    """measure_union(...).reference_rows() == the rows GetMask_Contours (:371-459) appended."""
    man, meas, _ = gold
    batch, (H, W) = {"union_dense": FX.union_dense, "blobs_rescaled": FX.blobs_rescaled,
                     "c1_maskrcnn": lambda: FX.c1_maskrcnn(golden_dir)}[tag]()
    info = man["measure"][tag]
    assert FX.digest(batch) == info["digest"], "fixture inputs changed: regenerate the goldens"
    exact = rows = 0
    for key, classes in [(f"cls{c}", [c]) for c in range(4)] + [("all", None)]:
        ut = uwcv.measure_union(batch, (H, W), classes_of_interest=classes)
        for k in range(len(batch)):
            want = meas[f"{tag}/img{k}.tif/{key}"]
            e, n = _close(ut.reference_rows(image_idx=k), want, (tag, k, key))
            exact += e * n
            rows += n
    for k, inst in enumerate(batch):
        ref = info["counts"][f"img{k}.tif"]
        c = uwcv.get_counts(inst)                 # intended histogram over ids 0..3
        # the reference's GetCounts compares against ids 1..4 and repeats ``== 3`` (:359-362)
        assert [c[1], c[2], c[3], c[3]] == [ref["SList"], ref["WTList"], ref["PTList"], ref["PList"]]
    assert rows > 20
    print(f"{tag}: {rows} reference rows matched, {exact} in bit-exact blocks")


def test_driver_loop_rows_and_shape_descriptor_csv(gold, tmp_path):
    """nn_inference.py:487-559 per keyword: union rows of all images of the folder -> moving
    average -> ShapeDescriptor.csv, the text the reference wrote."""
    man, meas, _ = gold
    batch, (H, W) = FX.blobs_rescaled()
    drv = man["measure"]["blobs_rescaled"]["driver"]
    for c, kw in enumerate(uwcv.CLASS_KEYWORDS):
        ut = uwcv.measure_union(batch, (H, W), classes_of_interest=[c])
        rows = ut.reference_rows()
        want = meas[f"blobs_rescaled/driver/{kw}"]
        _close(rows, want, kw)
        p = tmp_path / f"ShapeDescriptor_{kw}.csv"
        uwcv.write_shape_descriptor_csv(str(p), rows)
        text = p.read_text()
        if text != drv[kw]["shape_csv"]:
            # a 1-ulp difference in a float64 column may flip a 2-dp rounding: allow none in the
            # float32 / int-truncated columns, report any other
            a = [l.split(",") for l in text.split()]
            b = [l.split(",") for l in drv[kw]["shape_csv"].split()]
            assert len(a) == len(b)
            diff = [(i, j) for i in range(len(a)) for j in range(9) if a[i][j] != b[i][j]]
            assert all(uwcv.schema.CSV_KINDS[j] != "f32" for _, j in diff), diff
            assert len(diff) <= 1, diff


def test_postprocess_masks_equal_the_reference_function(gold):
    """uwcv.postprocess_masks == nn_inference.py:265-306 mask for mask (None / [] / truncation)."""
    man, _, cl = gold
    info = man["cleanup"]["postprocess_masks"]
    for case in FX.bool_mask_cases():
        H, W = case["masks"].shape[1:]
        want = info[case["name"]]
        got = uwcv.postprocess_masks(case["masks"], case["scores"], np.zeros((H, W, 3), np.uint8))
        if want["kind"] == "None":
            assert got is None, case["name"]
        elif want["kind"] == "empty":
            assert got == [], case["name"]
        else:
            ref = np.unpackbits(cl[f"pp/{case['name']}"], axis=-1)[..., :W]
            assert len(got) == ref.shape[0], case["name"]
            for k in range(len(got)):
                assert got[k].dtype == np.uint8 and np.array_equal(got[k], ref[k]), (case["name"], k)


def test_export_rle_equals_the_reference_csv(gold, tmp_path):
    """export_rle + write_rle_csv == the R50_flip_.csv the reference's export loop (:313-336) wrote."""
    man, _, cl = gold
    batch, names, (H, W) = FX.export_batch()
    assert FX.digest(batch) == man["cleanup"]["export"]["digest"]
    got = uwcv.export_rle(batch, (H, W), names)
    text = bytes(cl["export/csv"]).decode()
    lines = text.split("\n")[1:-1]
    assert got.image_id == man["cleanup"]["export"]["image_ids"]
    bad = [k for k, line in enumerate(lines) if line != f"{got.image_id[k]},{got.encoded_pixels[k]}"]
    assert not bad, f"{len(bad)} of {len(lines)} rows differ, first {bad[:5]}"
    p = tmp_path / "R50_flip_.csv"
    uwcv.write_rle_csv(str(p), got)
    assert p.read_text() == text
