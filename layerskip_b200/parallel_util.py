"""Host-side helpers for the multi-process (one process per GPU) paths.  Pure plumbing over
`torch.distributed`; covered on CPU by world_size-2 gloo tests."""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_prompts(prompts: Sequence, rank: int, world: int) -> List:
    """Replica mode: rank r serves prompts r, r+world, ... (disjoint, covering)."""
    return list(prompts[rank::world])


def broadcast_bytes(payload: bytes, n: int, src: int = 0, group=None, device=None) -> bytes:
    """Rank `src` provides `payload` (n bytes); every rank returns the same bytes."""
    import torch.distributed as dist
    if dist.get_rank(group) == src:
        assert len(payload) == n
        buf = torch.tensor(list(payload), dtype=torch.uint8)
    else:
        buf = torch.zeros(n, dtype=torch.uint8)
    if device is not None and dist.get_backend(group) == "nccl":
        buf = buf.to(device)
    root = dist.get_global_rank(group, src) if group is not None else src
    dist.broadcast(buf, src=root, group=group)
    return bytes(buf.cpu().tolist())


def reduce_scalar(x: float, op: str = "sum", group=None, device=None) -> float:
    """All-reduce one float (sum / max) — bench.py's max-over-ranks timing and token totals."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64)
    if device is not None and dist.get_backend(group) == "nccl":
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=group)
    return float(t.item())
