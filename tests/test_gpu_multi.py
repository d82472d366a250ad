"""Two-GPU tests (skipped on a single-GPU box): the fused all-gather of the measurement table
-- peer stores of the trace kernel into symmetric memory + a signal barrier -- equals the NCCL
all-gather and the single-process table."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "uw-com-vision_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import uwcv
    from uwcv import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_WORLD_SIZE"] = str(world)            # one node, as torchrun would say
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        H = W = 512
        n_images = 6
        mine = uwcv.shard_indices(n_images, rank, world)
        batches = []
        for rep in range(3):                                   # three calls: both table sets are reused
            batches.append([synth.blob_instances(b, 40 + 9 * b + rep, H, W, seed=60 + rep) for b in mine])
        eng = uwcv.Engine.get(dev)
        fused = eng.fused_gather() is not None
        outs = []
        stream = uwcv.MeasurementStream(dev, depth=2)
        offs = [rank]                                          # image index = position in the whole set
        for t in stream.map(batches, (H, W), gather=True):
            outs.append((t.ints.copy(), t.floats.copy()))
        # gather to rank 0 with the HOST as the sink (node-shared pinned table; every rank copies
        # its own rows), then with the device gather restricted to the destination's table
        host_ok = eng.host_table() is not None
        dst_outs = {}
        for sink in (("host",) if host_ok else ()) + ("device",):
            got = []
            for t in stream.map(batches, (H, W), gather=True, gather_dst=0, gather_sink=sink):
                got.append((t.ints.copy(), t.floats.copy()))
                del t
            dst_outs[sink] = got
        # the NCCL path on the same inputs
        os.environ["UWCV_NO_FUSED_GATHER"] = "1"
        eng._fused = None
        ref = [uwcv.measure_instances(b, (H, W), gather=True, device=dev) for b in batches]
        own = [uwcv.measure_instances(b, (H, W), device=dev, image_idx_offset=0) for b in batches]
        q.put((rank, fused, outs, [(r.ints.copy(), r.floats.copy()) for r in ref], host_ok, dst_outs,
               [(r.ints.copy(), r.floats.copy()) for r in own]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_gather_equals_nccl_all_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "symmetric-memory gather was not available on this box"
    for rank, fused, outs, ref, host_ok, dst_outs, own in res:
        for (gi, gf), (ri, rf) in zip(outs, ref):
            assert np.array_equal(gi, ri)
            assert np.array_equal(gf, rf, equal_nan=True)
        assert host_ok, "the node-shared host table was not available on this box"
        for sink, got in dst_outs.items():
            for k, (gi, gf) in enumerate(got):
                if rank == 0:                      # the destination holds the whole job's table
                    assert np.array_equal(gi, ref[k][0]), (sink, k)
                    assert np.array_equal(gf, ref[k][1], equal_nan=True), (sink, k)
                else:                              # the others their own rows (image index aside)
                    assert gi.shape == own[k][0].shape, (sink, k)
                    assert np.array_equal(gi[:, 1:], own[k][0][:, 1:]), (sink, k)
                    assert np.array_equal(gf, own[k][1], equal_nan=True), (sink, k)
    # both ranks hold the same whole-job table
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True)
