"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol,
argument validation (no compute), the schema, the box glue, grouping/report and the
world_size-2 gloo all-gather."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

import uwcv
from uwcv import _lib, api, dist as udist, grouping, schema
from uwcv.build import LIB_PATH
from oracle import d2, measure as M, pipeline as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "run __graft_entry__.build() first"
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "uwcv.h")).read()
    declared = set(re.findall(r"\b(uwcv_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.uwcv_version() >= 100
    assert L.uwcv_plane_row_words(2048) == 64 and L.uwcv_plane_row_words(1000) == 32
    assert L.uwcv_plane_row_words(33) == 4
    assert L.uwcv_workspace_bytes(10, 1000) >= 10 * 32 + 1000 * 28


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call (error behaviour of the ABI)."""
    L = _lib.lib()
    st = (C.c_int64 * 4)()
    call = lambda **k: L.uwcv_paste_measure(  # noqa: E731
        None, None, None, None, None, None, k.get("N", 1), k.get("H", 8), k.get("W", 8),
        k.get("thr", 0.5), k.get("ppm", 0.85), None, None, None, None, 0,
        k.get("status", C.addressof(st)), None)
    assert call(N=-1) == -2
    assert call(H=0) == -2
    assert call(W=40000) == -8
    assert call(thr=0.0) == -4
    assert call(thr=-1.0) == -4
    assert call(ppm=0.0) == -2
    assert call(status=None) == -1
    assert call() == -1                                   # N > 0 with NULL arrays
    assert _lib.strerror(-4).startswith("mask threshold")
    assert _lib.strerror(-7).startswith("tile words")
    assert L.uwcv_unpack_planes(None, 1, 8, 8, None, None) == -1
    assert L.uwcv_unpack_planes(None, 0, 8, 8, None, None) == 0
    # the plane allocator validates before it touches the driver; without a driver it reports a launch error
    ptr, comp = C.c_void_p(), C.c_int(7)
    assert L.uwcv_planes_alloc(0, C.byref(ptr), C.byref(comp)) == -2 and comp.value == 0
    assert L.uwcv_planes_alloc(1 << 20, None, None) == -1
    assert L.uwcv_planes_free(None) == 0
    assert L.uwcv_planes_free(C.c_void_p(4096)) == -2                       # not one of ours
    if not torch.cuda.is_available():
        assert L.uwcv_planes_alloc(1 << 20, C.byref(ptr), None) == -6 and not ptr.value
    # split pipeline: stage 8 (planes from the tiles) without a plane buffer is refused before any launch
    fake = 4096                                          # non-NULL, 16-byte aligned, never dereferenced
    assert L.uwcv_paste_measure_stages(fake, fake, None, None, None, None, 1, 8, 8, 0.5, 0.85, None, fake, fake,
                                       fake, 1 << 20, C.addressof(st), None, 8 | 16) == -1
    off = (C.c_int64 * 2)(0, -5)
    assert L.uwcv_nms_filter(None, None, None, off, 1, 4, 0.5, 0.5, 10, None, C.addressof(st), None, 0, None) == -2
    off = (C.c_int64 * 2)(1, 5)
    assert L.uwcv_nms_filter(None, None, None, off, 1, 4, 0.5, 0.5, 10, None, C.addressof(st), None, 0, None) == -2


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from uwcv import synth
    batch = [synth.blob_instances(0, 5, 64, 64)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uwcv.measure_instances(batch, (64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uwcv.paste_masks_in_image(torch.rand(1, 28, 28), torch.tensor([[0., 0., 5., 5.]]), (8, 8))


def test_schema_matches_oracle():
    assert list(schema.INT_COLUMNS) == M.INT_COLUMNS
    assert list(schema.FLOAT_COLUMNS) == M.FLOAT_COLUMNS
    assert list(schema.CSV_COLUMNS) == M.CSV_COLUMNS
    assert list(schema.CLASS_NAMES) == M.CLASS_NAMES
    assert list(schema.CLASS_KEYWORDS) == M.KEYWORDS
    assert schema.NUM_INT == 20 and schema.NUM_FLOAT == 30
    # enum order in the CUDA header matches the schema
    cu = open(os.path.join(ROOT, "uw-com-vision_b200", "csrc", "uwcv_common.cuh")).read()
    enum_i = re.search(r"enum IntCol \{(.*?)\};", cu, re.S).group(1)
    names_i = [n.strip() for n in enum_i.replace("= 0", "").split(",") if n.strip()]
    assert len(names_i) == 20 and names_i[5] == "I_AREA" and names_i[10] == "I_M10"
    enum_f = re.search(r"enum FloatCol \{(.*?)\};", cu, re.S).group(1)
    names_f = [n.strip() for n in enum_f.replace("= 0", "").split(",") if n.strip()]
    assert len(names_f) == 30 and names_f[14] == "F_CAREA" and names_f[21] == "F_FERET"


def test_box_glue_equals_detectron2_restatement():
    torch.manual_seed(1)
    b = torch.rand(200, 4) * 300 - 20
    b[:, 2:] = b[:, :2] + (torch.rand(200, 2) - 0.2) * 80
    inst = d2.Instances((256, 333), pred_boxes=d2.Boxes(b.clone()), scores=torch.rand(200),
                        pred_classes=torch.zeros(200, dtype=torch.int64),
                        pred_masks=torch.rand(200, 1, 28, 28))
    res = P.postprocess_boxes(inst, (320, 416))
    mine, keep = api.scale_clip_boxes(b, (256, 333), (320, 416))
    assert np.array_equal(mine[keep].numpy(), res.pred_boxes.tensor.numpy())
    assert int(keep.sum()) == len(res)
    with pytest.raises(AssertionError):
        api.scale_clip_boxes(torch.tensor([[0., 0., float("nan"), 1.]]), (10, 10), (10, 10))


def test_tile_words_mirror():
    b = torch.tensor([[10., 10., 50., 40.], [0., 0., 0., 5.], [100., 100., 4000., 4000.],
                      [-5., -5., 3., 3.]])
    w = api.tile_words(b, 256, 256)
    assert w > 0
    assert api.tile_words(torch.zeros(0, 4), 64, 64) == 0
    assert api.tile_words(torch.tensor([[0., 0., 0., 5.]]), 64, 64) == 0     # empty box: no tile


def test_instances_container():
    inst = uwcv.Instances((10, 20), pred_boxes=uwcv.Boxes(torch.zeros(3, 4)),
                          scores=torch.tensor([.1, .2, .3]))
    assert len(inst) == 3 and inst.image_size == (10, 20)
    assert inst._fields["scores"] is inst.scores          # nn_inference.py:326-327 access path
    assert len(inst[1]) == 1 and len(inst[torch.tensor([True, False, True])]) == 2
    with pytest.raises(ValueError):
        inst.set("bad", torch.zeros(2))
    with pytest.raises(AttributeError):
        inst.nope
    assert inst.has("scores") and not inst.has("pred_masks")
    assert uwcv.get_counts(uwcv.Instances((1, 1), pred_classes=torch.tensor([0, 3, 3, 1]))) == [1, 1, 0, 2]


def test_grouping_and_report(tmp_path):
    rng = np.random.default_rng(0)
    n = 50
    ints = np.zeros((n, schema.NUM_INT), np.int64)
    floats = rng.random((n, schema.NUM_FLOAT)) * 300
    ints[:, schema.ICOL["class_id"]] = rng.integers(0, 4, n)
    ints[:, schema.ICOL["valid"]] = 1
    ints[:5, schema.ICOL["valid"]] = 0
    t = uwcv.MeasurementTable(ints, floats)
    recs = uwcv.group_by_class(t)
    assert [r["count"] for r in recs] == [int((ints[:, 2] == k).sum()) for k in range(4)]
    ok = (ints[:, 3] == 1) & (floats[:, schema.FCOL["contour_area"]] >= 100)
    assert sum(r["measured"] for r in recs) == int(ok.sum())
    rows = t.for_class(2).reference_rows()
    assert rows.shape[1] == 9
    # report layer equals the oracle restatement of nn_inference.py:500-570
    sm, h = uwcv.report_class(rows)
    sm_o, h_o = M.report_class(rows)
    assert np.array_equal(sm, sm_o)
    for k in h_o:
        assert np.array_equal(h[k][0], h_o[k][0]) and np.array_equal(h[k][1], h_o[k][1])
    assert uwcv.moving_average([1, 2, 3, 4.005]) == M.moving_average([1, 2, 3, 4.005])
    p = tmp_path / "ResultsPore_.csv"
    uwcv.write_results_csv(str(p), sm)
    import pandas as pd
    df = pd.read_csv(p, index_col=0)
    assert list(df.columns) == list(schema.CSV_COLUMNS) and len(df) == sm.shape[0]
    uwcv.write_classes_csv(str(tmp_path / "classes.csv"), t)
    df = pd.read_csv(tmp_path / "classes.csv")
    assert list(df["class_name"]) == list(schema.CLASS_NAMES)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_images = 5
        mine = udist.shard_indices(n_images, rank, world)
        rows_i = torch.tensor([[b, j] + [0] * 18 for b in mine for j in range(b + 1)],
                              dtype=torch.int64).reshape(-1, 20)
        rows_f = rows_i[:, :1].double().repeat(1, 30) + 0.25
        gi, gf = udist.all_gather_table(rows_i, rows_f)
        gi, gf = udist.sort_rows(gi, gf)
        q.put((rank, gi.numpy(), gf.numpy()))
    finally:
        dist.destroy_process_group()


def test_all_gather_table_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.array([[b, j] for b in range(5) for j in range(b + 1)])
    for rank, gi, gf in outs:
        assert np.array_equal(gi[:, :2], expect)
        assert np.array_equal(gf[:, 0], expect[:, 0] + 0.25)
    assert udist.shard_indices(5, 0, 2) == [0, 2, 4] and udist.shard_indices(5, 1, 2) == [1, 3]


def test_tile_major_sharding_partitions_the_instances():
    """configs[3]-style partition of one micrograph: every instance lands on exactly one rank,
    whichever grid is used, and neighbours in space land together."""
    g = torch.Generator().manual_seed(4)
    H = W = 4096
    c = torch.rand((5000, 2), generator=g) * torch.tensor([W, H])
    wh = torch.rand((5000, 2), generator=g) * 40 + 8
    boxes = torch.cat((c - wh / 2, c + wh / 2), dim=1)
    boxes[0] = torch.tensor([-30.0, -30.0, 10.0, 10.0])             # centre outside the image: clamped
    boxes[1] = torch.tensor([W - 5.0, H - 5.0, W + 50.0, H + 50.0])
    for world, grid in ((1, None), (2, None), (4, None), (8, None), (8, (4, 4)), (3, None)):
        shards = [udist.shard_instances_by_tile(boxes, (H, W), r, world, grid) for r in range(world)]
        allidx = torch.cat(shards)
        assert allidx.numel() == 5000 and torch.equal(torch.sort(allidx)[0], torch.arange(5000))
        for s_ in shards:
            assert torch.equal(s_, torch.sort(s_)[0])
    s4 = [udist.shard_instances_by_tile(boxes, (H, W), r, 4) for r in range(4)]      # 2 x 2 grid
    own = torch.empty(5000, dtype=torch.long)
    for r, s_ in enumerate(s4):
        own[s_] = r
    cx, cy = (boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2
    inside = (cx > 0) & (cx < W) & (cy > 0) & (cy < H)
    want = (cy >= H / 2).long() * 2 + (cx >= W / 2).long()
    assert torch.equal(own[inside], want[inside])


def test_scale_clip_equals_detectron2_column_ops_and_words_bound():
    """The whole-tensor form of Boxes.scale / clip / nonempty is bit-identical to Detectron2's
    column-wise ops (oracle restatement), non-finite boxes raise as Boxes.clip does, and the cheap
    workspace bound never undercuts the exact tile-word count."""
    from oracle import d2
    from uwcv import api
    g = torch.Generator().manual_seed(0)
    for in_size, out_size in (((800, 1333), (1024, 1024)), ((512, 512), (2048, 2048)),
                              ((1000, 700), (333, 777)), ((640, 640), (640, 640)),
                              ((100, 200), (300, 600))):
        bx = (torch.rand((4000, 4), generator=g) - 0.2) * 1500
        ref = d2.Boxes(bx.clone())
        ref.scale(out_size[1] / in_size[1], out_size[0] / in_size[0])
        ref.clip(out_size)
        mine, keep = api.scale_clip_boxes(bx, in_size, out_size)
        assert torch.equal(mine, ref.tensor) and torch.equal(keep, ref.nonempty())
    for bad in (float("nan"), float("inf"), -float("inf")):
        bx = torch.rand((10, 4))
        bx[3, 1] = bad
        with pytest.raises(AssertionError):
            api.scale_clip_boxes(bx, (10, 10), (10, 10))
    for trial in range(100):
        c = torch.rand((60, 2), generator=g) * 300 - 50
        wh = torch.exp(torch.rand((60, 2), generator=g) * 9 - 3)
        bx, _ = api.scale_clip_boxes(torch.cat((c, c + wh), 1), (256, 256), (256, 256))
        assert api.tile_words_bound(bx) >= api.tile_words(bx, 256, 256)
    assert api.tile_words_bound(torch.zeros((0, 4))) == 0


def test_gather_struct_mirrors_the_header():
    """ctypes mirror of `uwcv_gather` (include/uwcv.h): 16 peers, pointer arrays after two
    32-bit fields and the 64-bit row base."""
    import ctypes as C
    import re
    from uwcv import _lib
    hdr = open(os.path.join(ROOT, "include", "uwcv.h")).read()
    peers = int(re.search(r"#define\s+UWCV_MAX_PEERS\s+(\d+)", hdr).group(1))
    assert peers == _lib.MAX_PEERS
    assert C.sizeof(_lib.Gather) == 4 + 4 + 8 + 2 * peers * 8
    assert _lib.Gather.row_base.offset == 8
    assert _lib.Gather.rows_i.offset == 16 and _lib.Gather.rows_f.offset == 16 + peers * 8


def test_release_library_reads_no_environment_knobs():
    """the sweep / elimination knobs (UWCV_DEBUG_SKIP skips the paste!) exist
    only in the -DUWCV_TUNING build; the shipped library does not even contain their names (the
    only getenv importer left is the static CUDA runtime) and exports no tuning entry point."""
    import subprocess
    blob = open(LIB_PATH, "rb").read()
    for knob in (b"UWCV_DEBUG_SKIP", b"UWCV_FILL", b"UWCV_ZERO_KB", b"UWCV_PASTE_ROT", b"UWCV_BAND_KB",
                 b"UWCV_PASTE_CTAS", b"UWCV_TILE_PER_CTA", b"UWCV_TRACE_LANES"):
        assert knob not in blob, knob
    assert "uwcv_tuning_" not in subprocess.run(["nm", "-D", "--defined-only", LIB_PATH],
                                                capture_output=True, text=True).stdout


def test_host_lease_releases_when_the_last_view_dies():
    """Explicit ownership of returned host buffers (no sys.getrefcount): the numpy
    arrays handed out hang off a HostLease; the release callback runs once, when the table AND every
    slice taken from it are gone."""
    import gc
    from uwcv.api import HostLease
    buf = torch.arange(100, dtype=torch.int64)
    fired = []
    arr = HostLease(buf.data_ptr(), 100, lambda: fired.append(1), keep=buf).array()
    rows = arr[10:30].reshape(2, 10)
    col = rows[:, 3]
    as_f = arr[40:50].view(np.float64)
    assert rows[1, 2] == 22
    del arr, rows
    gc.collect()
    assert not fired                      # two views are still alive
    del col
    gc.collect()
    assert not fired
    del as_f
    gc.collect()
    assert fired == [1]
