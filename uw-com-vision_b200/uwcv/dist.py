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
