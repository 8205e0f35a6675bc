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


# ---- demodulation chain: packet-aligned segments with a warm-up halo ---------------------------------
from rfanalyzer_b200.sharding import ShardedChain, default_halo_packets, shard_packets


def test_shard_packets_cover_the_recording_on_packet_boundaries():
    for total, packet in ((0, 8192), (1, 8192), (8192 * 7, 8192), (8192 * 7 + 5, 8192), (65536 * 33 - 1, 65536)):
        for world in (1, 2, 3, 8):
            segs = [shard_packets(total, packet, world, r) for r in range(world)]
            assert segs[0][0] == 0 and sum(n for _, n in segs) == total
            for (a, n), (b, m) in zip(segs, segs[1:]):
                assert a + n == b or m == 0
                assert b % packet == 0 or m == 0


class _FakePlan:
    """Counts what ShardedChain asks of a chain plan: 1 audio sample per 50 input samples."""
    class desc:
        packet_samples, mode, format = 1000, 3, 1

    def __init__(self):
        self.calls, self.pos = [], 0

    def seek(self, n):
        self.pos = n
        self.calls.append(("seek", n))
        return n // 50

    def max_audio(self, n):
        return n // 50 + 64

    def process(self, iq, n, audio):
        assert len(iq) == 2 * n
        self.calls.append(("process", self.pos, n))
        self.pos += n
        return n // 50


def _chain_worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = _FakePlan()
    sc = ShardedChain(plan, rank, world, halo_packets=2)
    halo_start, first, n = sc.segment(total)
    iq = np.zeros(2 * (first + n - halo_start), np.uint8)
    audio = np.zeros(plan.max_audio(max(first - halo_start, n)), np.float32)
    index, got = sc.process(iq, total, audio)
    layout = sc.gather_layout(index, got)
    q.put((rank, plan.calls, layout))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_chain_layout_over_gloo():
    world, total = 2, 1000 * 9 + 400
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chain_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    (r0, calls0, lay0), (r1, calls1, lay1) = res
    assert lay0 == lay1                                   # every rank knows every piece
    assert lay0[0][0] == 0 and lay0[0][0] + lay0[0][1] == lay0[1][0]   # contiguous audio
    assert lay0[1][0] + lay0[1][1] == total // 50
    assert calls0 == [("seek", 0), ("process", 0, 5000)]  # rank 0: no halo
    assert calls1 == [("seek", 3000), ("process", 3000, 2000), ("process", 5000, 4400)]
    assert default_halo_packets(3) == 1 and default_halo_packets(1) == 256
