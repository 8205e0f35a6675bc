"""Host-side logic of the multi-GPU path (rfanalyzer_b200/sharding.py) over gloo, world_size 2:
segment partitioning, peak MAX all-reduce, and assembly of the newest L+1 rows across ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rfanalyzer_b200.sharding import assemble_tail, shard_frames, tail_owner_plan


def test_shard_frames_cover_everything():
    for total in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            segs = [shard_frames(total, world, r) for r in range(world)]
            assert segs[0][0] == 0 and sum(n for _, n in segs) == total
            for (a, n), (b, _) in zip(segs, segs[1:]):
                assert a + n == b


def test_tail_owner_plan():
    assert tail_owner_plan(4096, 8, 8) == [(7, 9)]
    assert tail_owner_plan(10, 4, 8) == [(3, 2), (2, 2), (1, 3), (0, 2)]
    assert tail_owner_plan(3, 2, 8) == [(1, 1), (0, 2)]


def _worker(rank, world, port, total_frames, L, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)
    rows_all = (rng.standard_normal((total_frames, n)) * 10 - 60).astype(np.float32)
    first, nloc = shard_frames(total_frames, world, rank)
    local = torch.from_numpy(rows_all[first:first + nloc])
    peaks = local.max(dim=0).values if nloc else torch.full((n,), -999999.0)
    dist.all_reduce(peaks, op=dist.ReduceOp.MAX)
    mine = torch.full((L + 1, n), -9999.0)
    take = min(L + 1, nloc)
    if take:
        mine[:take] = torch.flip(local[nloc - take:], dims=[0])
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    tail = assemble_tail(gathered, total_frames, world, L)
    q.put((rank, peaks.numpy(), tail.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_frames,L", [(64, 8), (5, 8), (13, 3)])
def test_two_rank_reduction_over_gloo(total_frames, L):
    world, n = 2, 32
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_frames, L, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(99)
    rows_all = (rng.standard_normal((total_frames, n)) * 10 - 60).astype(np.float32)
    want_tail = np.full((L + 1, n), -9999.0, np.float32)
    k = min(L + 1, total_frames)
    want_tail[:k] = rows_all[::-1][:k]
    for _, peaks, tail in results:
        assert np.array_equal(peaks, rows_all.max(axis=0))
        assert np.array_equal(tail, want_tail)
