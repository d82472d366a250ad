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
