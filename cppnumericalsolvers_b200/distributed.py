"""Multi-GPU plumbing: one process per GPU, the batch sharded embarrassingly.

The path has no data-path collective (instances are independent: solver history
and stopping state are per solve, lbfgs.h:306-323 / solver.h:184).  The only
exchange is the global stop test: each rank packs its per-instance status into a
convergence bitmap and ONE all-gather makes every rank see all of them.
Backend: NCCL on GPUs (NVLink/NVSwitch), gloo in the CPU tests of this logic.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous instance range [lo, hi) owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_done_bitmap(status: torch.Tensor) -> torch.Tensor:
    """Host/any-device twin of cno_done_bitmap: bit i of word i//32 = instance i
    has terminated (status is neither Continue nor NotStarted). int32 words."""
    done = (status != 0) & (status != -1)
    n = done.numel()
    pad = (-n) % 32
    if pad:
        done = torch.cat([done, done.new_zeros(pad)])
    bits = done.view(-1, 32).to(torch.int64)
    weights = (1 << torch.arange(32, dtype=torch.int64, device=bits.device))
    words = (bits * weights).sum(1)
    words = torch.where(words >= (1 << 31), words - (1 << 32), words)
    return words.to(torch.int32)


def gather_done_bitmaps(local_words: torch.Tensor, max_words: int | None = None) -> torch.Tensor:
    """ONE all-gather of the per-rank convergence bitmaps -> [world, max_words]
    (shards may differ by one word; shorter ones are zero padded)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_words.unsqueeze(0)
    if max_words is None:
        n = torch.tensor([local_words.numel()], dtype=torch.int64, device=local_words.device)
        dist.all_reduce(n, op=dist.ReduceOp.MAX)
        max_words = int(n.item())
    buf = local_words.new_zeros(max_words)
    buf[: local_words.numel()] = local_words
    out = local_words.new_empty(world * max_words)
    dist.all_gather_into_tensor(out, buf)
    return out.view(world, max_words)


def all_done(gathered: torch.Tensor, global_batch: int) -> bool:
    """Global stop test: every instance of every shard has a terminal status."""
    world = gathered.shape[0]
    total = 0
    for r in range(world):
        lo, hi = shard_range(global_batch, r, world)
        n = hi - lo
        w = gathered[r].to(torch.int64) & 0xFFFFFFFF
        full, tail = divmod(n, 32)
        cnt = 0
        for k in range(32):
            cnt += int(((w[:full] >> k) & 1).sum().item())
        if tail:
            cnt += bin(int(w[full].item()) & ((1 << tail) - 1)).count("1")
        total += cnt
    return total == global_batch
