"""Contact exchange for a world whose pairs are sharded across ranks.

One all-gather of per-rank counts, then ONE all-gather of fixed-stride padded record blocks — the
only collective on the path.  Works with any torch.distributed backend (nccl on the GPU box, gloo
in the CPU tests); the tensors are plain uint8 views of pk_contact records (88 bytes each).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_BYTES = 88


def allgather_records(local: torch.Tensor, count: int, record_bytes: int = RECORD_BYTES):
    """local: uint8 tensor holding `count` records (may be longer).  Returns (gathered uint8 tensor of
    total·record_bytes bytes in rank order, per-rank counts list)."""
    world = dist.get_world_size()
    dev = local.device
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    stride = max(max(counts), 1) * record_bytes
    send = torch.zeros(stride, dtype=torch.uint8, device=dev)
    send[: count * record_bytes] = local[: count * record_bytes]
    recv = torch.empty(world * stride, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send)
    parts = [recv[r * stride : r * stride + counts[r] * record_bytes] for r in range(world)]
    return torch.cat(parts), counts
