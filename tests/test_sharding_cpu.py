"""Host-side logic of the multi-GPU modes on CPU: agent partitioning (population mode, no collective on the data
path) and the data-parallel batch split / gradient exchange, with a 2-rank gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sac.population import shard_agents, split_batch


def test_shard_agents_partitions_exactly():
    for n, w in [(1024, 8), (1024, 3), (5, 4), (7, 7), (10, 1)]:
        got = [list(shard_agents(n, w, r)) for r in range(w)]
        flat = [a for g in got for a in g]
        assert flat == list(range(n))
        assert max(map(len, got)) - min(map(len, got)) <= 1
    with pytest.raises(ValueError):
        shard_agents(4, 2, 2)


def test_split_batch():
    assert split_batch(65536, 8) == 8192 and split_batch(65536, 2) == 32768
    with pytest.raises(ValueError):
        split_batch(100, 3)


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
    try:
        # population mode: every rank owns a disjoint block; the only exchange is the metrics gather
        mine = list(shard_agents(11, world, rank))
        sizes = [len(shard_agents(11, world, r)) for r in range(world)]
        pad = torch.zeros(max(sizes), dtype=torch.float64)
        pad[: len(mine)] = torch.tensor([100.0 + a for a in mine], dtype=torch.float64)
        gathered = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        merged = np.concatenate([g.numpy()[:n] for g, n in zip(gathered, sizes)])
        # data-parallel mode: per-rank gradient shares carry 1/B_global, all-reduce(sum) gives the global mean gradient
        B, w = 64, world
        rng = np.random.default_rng(0)
        per_row = rng.standard_normal((B, 5))
        rows = slice(rank * split_batch(B, w), (rank + 1) * split_batch(B, w))
        share = torch.tensor(per_row[rows].sum(0) / B)
        dist.all_reduce(share, op=dist.ReduceOp.SUM)
        out.put((rank, merged.tolist(), share.numpy().tolist(), per_row.mean(0).tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_population_gather_and_dp_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, merged, share, mean in res:
        assert merged == [100.0 + a for a in range(11)]
        assert np.allclose(share, mean, rtol=1e-12, atol=1e-12)
