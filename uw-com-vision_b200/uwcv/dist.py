"""Image sharding and the single collective of the path: an all-gather of the
fixed-width measurement table.

One process per GPU (torchrun); rank r measures images ``b % world == r``.  Paste,
reduction, border trace and NMS need no exchange (every instance is confined to its own
box).  The only collective is ``all_gather_into_tensor`` of the int64 and float64 row
tables, padded to the largest per-rank row count (counts are gathered first, 8 bytes a
rank).  NCCL over NVLink on the GPU box; the same code runs on gloo for the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


def shard_instances_by_tile(boxes: torch.Tensor, image_size: Tuple[int, int], rank: int, world: int,
                            grid: Tuple[int, int] = None) -> torch.Tensor:
    """Tile-major partition of ONE huge micrograph (BASELINE configs[3]): the
    image is cut into a ``grid`` of gy x gx tiles (default: the most square factorisation of
    ``world``), tile t belongs to rank ``t % world`` and an instance belongs to the tile holding
    its box centre.  No halo is needed: an instance's mask is confined to its own box, so the rank
    that owns the centre measures the whole instance.  Returns the indices (int64, ascending) of
    the instances of ``rank``; the shards of all ranks are a partition of ``range(N)``.
    ``boxes``: N x 4 XYXY in ``image_size`` = (H, W) coordinates."""
    H, W = int(image_size[0]), int(image_size[1])
    if grid is None:
        gy = int(world ** 0.5)
        while world % gy:
            gy -= 1
        grid = (gy, world // gy)
    gy, gx = int(grid[0]), int(grid[1])
    b = boxes.detach().to(torch.float64)
    cx = ((b[:, 0] + b[:, 2]) * 0.5).clamp(0, W - 1e-9)
    cy = ((b[:, 1] + b[:, 3]) * 0.5).clamp(0, H - 1e-9)
    tx = torch.floor(cx * gx / W).long().clamp(0, gx - 1)
    ty = torch.floor(cy * gy / H).long().clamp(0, gy - 1)
    owner = (ty * gx + tx) % world
    return torch.nonzero(owner == rank, as_tuple=False).flatten()


def take_instances(inst, idx: torch.Tensor):
    """The sub-``Instances`` at ``idx`` with a field ``orig_idx`` holding the original positions
    (reported by ``measure_instances`` as the ``inst_idx`` column)."""
    from .structures import Boxes, Instances
    out = Instances(tuple(inst.image_size))
    fields = inst.get_fields() if hasattr(inst, "get_fields") else inst._fields
    for k, v in fields.items():
        if k == "orig_idx":
            continue
        if hasattr(v, "tensor"):
            out.set(k, Boxes(v.tensor[idx.to(v.tensor.device)]))
        else:
            out.set(k, v[idx.to(v.device)])
    base = fields["orig_idx"] if "orig_idx" in fields else torch.arange(len(inst))
    out.set("orig_idx", base[idx.to(base.device)].to(torch.int32))
    return out


def all_gather_table(rows_i: torch.Tensor, rows_f: torch.Tensor, group=None, counts=None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """rows_i [r, Ci] int64, rows_f [r, Cf] float64 on this rank's device -> the
    concatenation over ranks (rank order), with the padding rows removed.
    ``counts`` (list of per-rank row counts), when the caller already knows them, skips
    the count exchange and its host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows_i, rows_f
    world = dist.get_world_size(group)
    dev = rows_i.device
    if counts is None:
        cnt = torch.tensor([rows_i.shape[0]], dtype=torch.int64, device=dev)
        cts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cts, cnt, group=group)
        counts_h = cts.cpu().tolist()
    else:
        counts_h = [int(c) for c in counts]
    mx = max(max(counts_h), 1)
    ci, cf = rows_i.shape[1], rows_f.shape[1]
    pi = torch.zeros((mx, ci), dtype=torch.int64, device=dev)
    pf = torch.zeros((mx, cf), dtype=torch.float64, device=dev)
    pi[: rows_i.shape[0]] = rows_i
    pf[: rows_f.shape[0]] = rows_f
    gi = torch.empty((world * mx, ci), dtype=torch.int64, device=dev)
    gf = torch.empty((world * mx, cf), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gi, pi, group=group)
    dist.all_gather_into_tensor(gf, pf, group=group)
    if all(c == mx for c in counts_h):
        return gi, gf
    keep = torch.cat([torch.arange(r * mx, r * mx + c, device=dev) for r, c in enumerate(counts_h)])
    return gi[keep], gf[keep]


def sort_rows(rows_i: torch.Tensor, rows_f: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Canonical (image_idx, inst_idx) order, so that the N-GPU table equals the 1-GPU one."""
    key = rows_i[:, 0] * (1 << 32) + rows_i[:, 1]
    order = torch.argsort(key, stable=True)
    return rows_i[order], rows_f[order]


class FusedGather:
    """All-gather of the measurement rows WITHOUT a collective kernel: the tables of all ranks
    live in symmetric memory (``torch.distributed._symmetric_memory``: every rank maps every
    peer's buffer over NVLink), the border-trace kernel of each rank stores its finished rows
    into all of them (``uwcv_paste_measure_gather``), and a signal barrier on the same stream
    makes them complete.  Two table sets are used in turn so that the device->host read of call
    i never races the peers' stores of call i + 2.

    Construction is a collective (rendezvous); it raises when symmetric memory is not available
    for the group, and the caller then stays on ``all_gather_table`` (NCCL)."""

    def __init__(self, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        from ._lib import MAX_PEERS
        if self.world > MAX_PEERS:
            raise RuntimeError(f"FusedGather supports up to {MAX_PEERS} ranks")
        self.cap = 0
        self.sets = []                 # [(buffer, handle)] x 2
        self.read_done = [None, None]  # event: the last device->host read of the set finished
        self.parity = 0

    def _offsets(self, cap: int):
        from .schema import NUM_INT
        off_f = (cap * NUM_INT * 8 + 255) // 256 * 256
        return off_f

    def ensure(self, total_rows: int) -> None:
        """(Collective when it grows.)  Capacity for ``total_rows`` rows in both table sets."""
        from .schema import NUM_FLOAT
        if total_rows <= self.cap:
            return
        torch.cuda.synchronize(self.device)
        cap = max(int(total_rows * 1.1) + 64, 1024)
        nbytes = self._offsets(cap) + cap * NUM_FLOAT * 8 + 256
        self.sets = []
        for _ in range(2):
            buf = self._symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            hdl = self._symm.rendezvous(buf, self.group)
            self.sets.append((buf, hdl))
        self.cap = cap
        self.read_done = [None, None]

    def begin(self, counts, dst=None):
        """-> (ctypes uwcv_gather for this rank, set index, total rows).  ``dst``: only rank
        ``dst``'s table receives the rows (a gather to one rank: 1 / world of the NVLink stores)."""
        import ctypes as C
        from ._lib import Gather
        from .schema import NUM_FLOAT, NUM_INT  # noqa: F401
        counts = [int(c) for c in counts]
        total = sum(counts)
        self.ensure(total)
        k = self.parity
        self.parity ^= 1
        buf, hdl = self.sets[k]
        g = Gather()
        g.world = self.world
        g.dst_plus_1 = 0 if dst is None else int(dst) + 1
        g.row_base = sum(counts[:self.rank])
        off_f = self._offsets(self.cap)
        ptrs = list(hdl.buffer_ptrs)
        for p in range(self.world):
            g.rows_i[p] = int(ptrs[p])
            g.rows_f[p] = int(ptrs[p]) + off_f
        return g, k, total

    def barrier(self, k: int) -> None:
        """On the current stream, behind the trace kernel: all ranks' rows of set k have landed
        everywhere when it completes.  The rank first makes sure its own read of the OTHER set
        is over: passing this barrier is what lets the peers write that set again."""
        cur = torch.cuda.current_stream(self.device)
        other = self.read_done[k ^ 1]
        if other is not None:
            cur.wait_event(other)
        self.sets[k][1].barrier(channel=0)

    def tables(self, k: int, total: int):
        from .schema import NUM_FLOAT, NUM_INT
        buf = self.sets[k][0]
        off_f = self._offsets(self.cap)
        ti = buf[: self.cap * NUM_INT * 8].view(torch.int64).view(self.cap, NUM_INT)[:total]
        tf = buf[off_f: off_f + self.cap * NUM_FLOAT * 8].view(torch.float64).view(self.cap, NUM_FLOAT)[:total]
        return ti, tf


class SharedTableUnavailable(RuntimeError):
    """Raised on EVERY rank together when the node-shared table cannot be created or grown."""


class SharedHostTable:
    """The whole job's measurement table in ONE block of host memory shared by the ranks of a
    node: every rank copies its OWN rows (device -> host, 400 B per instance) to its row offset
    of a POSIX shared-memory segment that all ranks map and register with the CUDA driver
    (``cudaHostRegister``), so the table is complete on the host without any rank reading the
    other ranks' rows over PCIe and without a device-side collective: the host is the sink.

    Replaces, for the end-to-end path, "all-gather on the devices, then rank 0 copies the whole
    table" (8 ranks: 205 MB device->host through one root port per step, next to 1.6 GB of
    host->device mask traffic).  ``SETS`` tables are used in turn.  Completion is signalled by the
    copy engine itself: behind its rows every rank copies an 8-byte sequence number into a flag
    word of the segment (copies of one stream land in order), which the destination rank polls;
    a set is written again only after the reader of its previous table has let go of it
    (``HostLease`` on the returned arrays).  Construction / growth is a collective."""

    SETS = 4
    HEADER = 4096

    def __init__(self, device: torch.device, group=None):
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.cap = 0
        self.seq = 0
        self.mm = None
        self.buf = None
        self.gen = 0
        self.own_released = [0] * self.SETS
        self.flag_dev = torch.zeros(1, dtype=torch.int64, device=device)

    # -- layout ---------------------------------------------------------------------------
    def _set_bytes(self, cap: int):
        from .schema import NUM_FLOAT, NUM_INT
        bi = (cap * NUM_INT * 8 + 4095) // 4096 * 4096
        bf = (cap * NUM_FLOAT * 8 + 4095) // 4096 * 4096
        return bi, bf

    def ensure(self, total_rows: int) -> None:
        import mmap
        import os
        if total_rows <= self.cap:
            return
        torch.cuda.synchronize(self.device)
        self._close()
        cap = max(int(total_rows * 1.1) + 64, 1024)
        bi, bf = self._set_bytes(cap)
        nbytes = self.HEADER + self.SETS * (bi + bf)
        self.gen += 1
        name = [f"/dev/shm/uwcv_table_{os.getuid()}_{os.getpid()}_{self.gen}"]
        dist.broadcast_object_list(name, src=dist.get_global_rank(self.group, 0), group=self.group)
        path = name[0]
        # a segment that does not fit into /dev/shm would only fail when its pages are touched
        # (SIGBUS): every rank checks the free space first and all ranks give up together
        try:
            st = os.statvfs("/dev/shm")
            room = st.f_bavail * st.f_frsize >= nbytes + (64 << 20)
        except OSError:
            room = False
        ok = torch.tensor([1 if room else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            raise SharedTableUnavailable(f"/dev/shm has no room for a {nbytes >> 20} MiB shared table")
        if self.rank == 0:
            fd = os.open(path, os.O_CREAT | os.O_RDWR | os.O_EXCL, 0o600)
            os.ftruncate(fd, nbytes)
        dist.barrier(self.group)
        if self.rank != 0:
            fd = os.open(path, os.O_RDWR)
        self.mm = mmap.mmap(fd, nbytes)
        os.close(fd)
        dist.barrier(self.group)
        if self.rank == 0:
            os.unlink(path)                       # the mappings keep the segment alive
        self.buf = torch.frombuffer(self.mm, dtype=torch.uint8)
        rc = int(torch.cuda.cudart().cudaHostRegister(self.buf.data_ptr(), nbytes, 1))   # portable
        self.registered = rc == 0
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # all ranks or none
        if int(ok.item()) == 0:
            self._close()
            raise SharedTableUnavailable(f"cudaHostRegister of the shared table failed on some rank (here: {rc})")
        self.cap = cap
        self.hdr = self.buf[: self.HEADER].view(torch.int64)
        self.hdr_np = self.hdr.numpy()
        if self.rank == 0:
            self.hdr_np[:] = 0
        self.own_released = [0] * self.SETS
        self.seq = 0
        dist.barrier(self.group)

    def _close(self):
        if self.buf is not None and getattr(self, "registered", False):
            try:
                torch.cuda.cudart().cudaHostUnregister(self.buf.data_ptr())
            except Exception:                      # noqa: BLE001
                pass
        self.registered = False
        self.buf = None
        self.hdr = self.hdr_np = None
        self.mm = None                             # (unmapped when the last view dies)
        self.cap = 0

    def _views(self, k: int):
        from .schema import NUM_FLOAT, NUM_INT
        bi, bf = self._set_bytes(self.cap)
        o = self.HEADER + k * (bi + bf)
        ti = self.buf[o: o + self.cap * NUM_INT * 8].view(torch.int64).view(self.cap, NUM_INT)
        tf = self.buf[o + bi: o + bi + self.cap * NUM_FLOAT * 8].view(torch.float64).view(self.cap, NUM_FLOAT)
        return ti, tf, o, o + bi

    def _done_index(self, k: int, r: int) -> int:
        return k * self.world + r

    def _released_index(self, k: int) -> int:
        return 256 + k

    # -- one call --------------------------------------------------------------------------
    def begin(self, counts, timeout_s: float = 120.0):
        """-> (set, seq, first row of this rank, total rows).  Blocks (host) until the readers of
        the table that last used this set have released it."""
        import time
        counts = [int(c) for c in counts]
        total = sum(counts)
        self.ensure(total)
        seq = self.seq
        self.seq += 1
        k = seq % self.SETS
        need = seq + 1 - self.SETS                 # tables are numbered seq + 1 (0: never)
        t0 = time.time()
        while int(self.hdr_np[self._released_index(k)]) < need or self.own_released[k] < need:
            time.sleep(0)
            if time.time() - t0 > timeout_s:
                raise RuntimeError(
                    "uwcv: a MeasurementTable of an earlier host-gathered call is still referenced "
                    f"(set {k}); copy or drop such tables within {self.SETS - 2} calls")
        return k, seq, sum(counts[: self.rank]), total

    def copy_rows(self, k: int, seq: int, base: int, rows_i: torch.Tensor, rows_f: torch.Tensor):
        """On the current stream: this rank's rows -> its slice of set k, then the flag."""
        ti, tf, _, _ = self._views(k)
        n = int(rows_i.shape[0])
        if n:
            ti[base: base + n].copy_(rows_i, non_blocking=True)
            tf[base: base + n].copy_(rows_f, non_blocking=True)
        self.flag_dev.fill_(seq + 1)
        j = self._done_index(k, self.rank)
        self.hdr[j: j + 1].copy_(self.flag_dev, non_blocking=True)

    def wait_all(self, k: int, seq: int, timeout_s: float = 120.0) -> None:
        """(Host.)  Every rank's rows of table ``seq`` have landed in set k."""
        import time
        t0 = time.time()
        lo = self._done_index(k, 0)
        while int(self.hdr_np[lo: lo + self.world].min()) < seq + 1:
            time.sleep(0)
            if time.time() - t0 > timeout_s:
                raise RuntimeError("uwcv: timed out waiting for the other ranks' rows")

    def arrays(self, k: int, seq: int, lo: int, hi: int, whole: bool):
        """numpy views of rows [lo, hi) of set k hanging off a lease; when the last view dies the
        set is released (``whole``: in the shared header, for every rank; else for this rank)."""
        from .api import HostLease
        from .schema import NUM_FLOAT, NUM_INT
        hdr_np, idx, own = self.hdr_np, self._released_index(k), self.own_released

        def release():
            own[k] = max(own[k], seq + 1)
            if whole:
                hdr_np[idx] = max(int(hdr_np[idx]), seq + 1)

        _, _, oi, of = self._views(k)
        base = self.buf.data_ptr()
        lease = HostLease(base, self.buf.numel() // 8, release, keep=(self.buf, self.mm))
        arr = lease.array()
        r = hi - lo
        a = oi // 8 + lo * NUM_INT
        b = of // 8 + lo * NUM_FLOAT
        n_i = arr[a: a + r * NUM_INT].reshape(r, NUM_INT)
        n_f = arr[b: b + r * NUM_FLOAT].view("float64").reshape(r, NUM_FLOAT)
        return n_i, n_f

    def release_unread(self, k: int, seq: int, whole: bool) -> None:
        """A rank that hands out no view of set k for this call releases its part at once."""
        self.own_released[k] = max(self.own_released[k], seq + 1)
        if whole:
            i = self._released_index(k)
            self.hdr_np[i] = max(int(self.hdr_np[i]), seq + 1)
