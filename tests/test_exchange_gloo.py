"""N>1 host logic on CPU: the contact all-gather (world_size 2, gloo) and the shard arithmetic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from physkit_b200.exchange import allgather_records

    n = 3 + 4 * rank  # ragged: 3 and 7 records
    rec = np.zeros(n, dtype=[("key", np.uint64), ("rest", np.float64, 10)])
    rec["key"] = np.arange(n) + 1000 * rank
    rec["rest"] = rank + 0.5
    local = torch.from_numpy(rec.view(np.uint8).copy())
    got, counts = allgather_records(local, n)
    res = np.frombuffer(got.numpy().tobytes(), dtype=rec.dtype)
    ok = counts == [3, 7] and len(res) == 10 and list(res["key"]) == [0, 1, 2] + [1000 + i for i in range(7)]
    ok = ok and np.all(res["rest"][:3] == 0.5) and np.all(res["rest"][3:] == 1.5)
    # empty contribution from one rank
    got2, counts2 = allgather_records(local, 0 if rank == 0 else n)
    ok = ok and counts2 == [0, 7] and got2.numel() == 7 * 88
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_allgather_records_world2_gloo():
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()  # (no fork: the session may hold the host-emulation thread pool, tests/cpp/simt_host.h)
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_shard_ranges_cover_all_leaves():
    """Same arithmetic as pk_collide_resident: [m·r/c, m·(r+1)/c) partitions [0, m)."""
    for m in (0, 1, 2, 7, 1000, 1_000_003):
        for c in (1, 2, 3, 4, 8):
            edges = [m * r // c for r in range(c + 1)]
            assert edges[0] == 0 and edges[-1] == m
            assert all(edges[i] <= edges[i + 1] for i in range(c))
