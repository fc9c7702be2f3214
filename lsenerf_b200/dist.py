"""Multi-GPU plumbing: rays are sharded over ranks (each rank draws its own pixels, train.py:104), the only exchange
is one all-reduce(sum) per step of the parameter gradients over NCCL / NVLink (DDP-mean semantics of
lse_pipeline.py:96-97; the 1/W factor is folded into the fused Adam kernel's grad_scale)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist


def allreduce_grads(groups: Dict[str, list]) -> None:
    """Sum-all-reduce every existing .grad.  The 64 MiB hash-table gradient goes alone (no copy); the small tensors are
    flattened into one bucket so the step issues two collectives in total."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    big, small = [], []
    for ps in groups.values():
        for p in ps:
            if p.grad is None:
                continue
            (big if p.grad.numel() >= (1 << 20) else small).append(p.grad)
    handles = [dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True) for g in big]
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g)); off += g.numel()
    for h in handles:
        h.wait()
