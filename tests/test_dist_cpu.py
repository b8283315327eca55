"""world_size-2 gloo test (CPU) of the multi-rank host logic bench.py uses: scene partitioning without overlap or gaps,
throughput aggregation as total scenes over the SLOWEST rank's time."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from epnet_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = shard.scene_ids(17, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, ids)
    elapsed = 100.0 if rank == 0 else 250.0  # rank 1 is the slow one
    sps, ms = shard.aggregate_scenes_per_second(len(ids), elapsed)
    mx = shard.max_over_ranks([float(rank), 5.0 - rank])
    q.put((rank, gathered, sps, ms, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, sps, ms, mx in results:
        flat = [i for part in gathered for i in part]
        assert sorted(flat) == list(range(17)) and len(set(flat)) == 17
        assert abs(len(gathered[0]) - len(gathered[1])) <= 1
        assert ms == 250.0 and abs(sps - 17 / 0.25) < 1e-9
        assert mx == [1.0, 5.0]


def test_partition_edge_cases():
    assert shard.scene_ids(16, 8, 3) == [6, 7]
    assert shard.scene_ids(3, 8, 5) == []
    assert sum(len(shard.scene_ids(5, 4, r)) for r in range(4)) == 5
    assert shard.max_over_ranks([1.5]) == [1.5]  # not initialised: identity
