"""Contact exchange for a world whose pairs are sharded across ranks.

One all-gather of per-rank counts, then ONE all-gather of fixed-stride padded record blocks — the
only collective on the path.  Works with any torch.distributed backend (nccl on the GPU box, gloo
in the CPU tests); the tensors are plain uint8 views of pk_contact records (88 bytes each).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_BYTES = 88


_recv_cache = {}


def allgather_records(local: torch.Tensor, count: int, record_bytes: int = RECORD_BYTES, concat: bool = True):
    """local: uint8 tensor holding `count` records (may be longer).  Returns (gathered, per-rank counts list);
    gathered is one uint8 tensor of total·record_bytes bytes in rank order, or — concat=False — the list of
    per-rank views into the receive buffer (valid until the next call; spares a copy of everything)."""
    world = dist.get_world_size()
    dev = local.device
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=dev)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, cnt)
    counts = [int(c) for c in counts_t.tolist()]
    stride = max(max(counts), 1) * record_bytes
    if local.numel() >= stride:
        send = local[:stride]  # the padding beyond `count` records is never read back
    else:
        send = torch.zeros(stride, dtype=torch.uint8, device=dev)
        send[: count * record_bytes] = local[: count * record_bytes]
    key = (str(dev), world)
    recv = _recv_cache.get(key)
    if recv is None or recv.numel() < world * stride:
        recv = torch.empty(int(world * stride * 1.25) + 1024, dtype=torch.uint8, device=dev)
        _recv_cache[key] = recv
    dist.all_gather_into_tensor(recv[: world * stride], send.contiguous())
    parts = [recv[r * stride : r * stride + counts[r] * record_bytes] for r in range(world)]
    return (torch.cat(parts) if concat else parts), counts
