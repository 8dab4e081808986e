"""CPU suite (gloo, world_size 2): item sharding + the single variable-length gather used for N > 1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uniaudio2_b200.distributed import gather_variable, shard_indices


def test_shard_indices_partition():
    for n in (0, 1, 7, 8, 13):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in shard_indices(n, r, w))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        items = [torch.randint(0, 8192, (8, 5 + 3 * i), generator=g) for i in range(n_items)]  # same on every rank
        local = [items[i] for i in shard_indices(n_items, rank, world)]
        out = gather_variable(local, n_items, pad_value=-1)
        ok = len(out) == n_items and all(torch.equal(a, b) for a, b in zip(out, items))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [2, 5])
def test_gather_variable_gloo_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
