"""Host logic of the N > 1 path on CPU: frame sharding + ordered score gather over gloo (world size 2),
and the frame-selection semantics of the reference loop (turbo-metrics/src/lib.rs:385-400)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from turbo_metrics_b200.engine import Options, gather_scores, select_frames, shard_range


def test_shard_range_partitions_exactly():
    for n in [0, 1, 7, 300, 2400, 2401]:
        for world in [1, 2, 3, 4, 8]:
            parts = [shard_range(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_select_frames_matches_reference_loop():
    assert select_frames(10, 10, Options()) == [(i, i) for i in range(10)]
    assert select_frames(10, 10, Options(every=3)) == [(0, 0), (3, 3), (6, 6), (9, 9)]
    assert select_frames(10, 8, Options(skip=2, skip_ref=1)) == [(3, 2), (4, 3), (5, 4), (6, 5), (7, 6), (8, 7)]
    assert select_frames(10, 10, Options(skip_dis=4, frames=3)) == [(0, 4), (1, 5), (2, 6)]
    # `frames` bounds the decode count, not the number of scored pairs (lib.rs:390-398)
    assert select_frames(100, 100, Options(every=4, frames=10)) == [(0, 0), (4, 4), (8, 8)]
    assert select_frames(5, 3, Options()) == [(0, 0), (1, 1), (2, 2)]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = shard_range(n, rank, world)
    local = [100.0 - 0.25 * i - 1e-9 * i * i for i in shard]   # a stand-in score that encodes the frame index
    res = gather_scores(local, n, rank, world)
    if rank == 0:
        q.put(res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 300])
def test_ordered_gather_world2_gloo(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    res = q.get(timeout=120)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res == [100.0 - 0.25 * i - 1e-9 * i * i for i in range(n)]
