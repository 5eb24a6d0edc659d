"""CPU tier for the N>1 path: class sharding + the single all-gather of per-class scores, run with
world_size 2 and 3 on the gloo backend; plus the sharding arithmetic for every supported rank count."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from var_b200.scoring import gather_class_scores, shard_range


@pytest.mark.parametrize("K", [1000, 1001, 10, 7])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_range_partitions_classes(K, world):
    spans = [shard_range(K, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == K
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def _fake_score(labels: torch.Tensor) -> torch.Tensor:
    return torch.sin(labels.float() * 0.37) * 100 - labels.float() * 0.01


def _worker(rank: int, world: int, port: int, K: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(K, rank, world)
        local = _fake_score(torch.arange(lo, hi))
        full = gather_class_scores(local, K, rank, world)
        q.put((rank, full.clone(), int(torch.argmax(full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,K", [(2, 1000), (3, 1001), (2, 7)])
def test_class_sharded_scores_allgather_gloo(world, K):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_score(torch.arange(K))
    for rank, full, pred in res:
        assert torch.equal(full, ref), f"rank {rank}: gathered scores differ from the single-process result"
        assert pred == int(torch.argmax(ref))
