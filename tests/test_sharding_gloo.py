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


class _RowsPlan:
    """Stands in for SpectrumPlan on a CPU box: the "IQ" of a segment is its finished dB rows, so everything
    ShardedSpectrum does around the kernel -- segment ranges, local peak hold / average, the packed all-reduce, the
    straddling-tail gather and the average_rows call -- runs exactly as on the GPU."""

    def __init__(self, n, L):
        self.fft_size, self.avg_len, self.ctx = n, L, self

    @staticmethod
    def average_rows(rows, newest, direction, ring_rows, row_stride, valid, avg_len, n, avg):
        # AnalyzerSurface.kt:710-714: newest -> oldest, float32, missing rows count as -9999f
        s = torch.zeros(n, dtype=torch.float32)
        for r in range(avg_len + 1):
            s = s + (rows[newest + r * direction, :n] if r < valid else torch.full((n,), -9999.0))
        avg.copy_(s / np.float32(avg_len + 1))

    def process(self, iq, nframes, rows=None, peaks=None, avg=None, peaks_accumulate=False):
        if nframes == 0:
            return
        rows[:nframes] = iq[:nframes]
        m = rows[:nframes].max(dim=0).values
        peaks.copy_(torch.maximum(peaks, m) if peaks_accumulate else m)
        self.average_rows(rows, nframes - 1, -1, 0, self.fft_size, nframes, self.avg_len, self.fft_size, avg)


CASES = [(64, 8), (5, 8), (13, 3), (18, 8), (2, 0), (1, 4)]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for total_frames, L in CASES:          # one process group serves every case (spawning costs seconds)
        _one_case(rank, world, total_frames, L, n, q)
    dist.barrier()
    dist.destroy_process_group()


def _one_case(rank, world, total_frames, L, n, q):
    from rfanalyzer_b200.sharding import ShardedSpectrum
    rng = np.random.default_rng(99)
    rows_all = (rng.standard_normal((total_frames, n)) * 10 - 60).astype(np.float32)
    shard = ShardedSpectrum(_RowsPlan(n, L), rank, world)
    first, nloc = shard.local_range(total_frames)
    local = torch.from_numpy(rows_all[first:first + nloc].copy())
    rows = torch.zeros((max(nloc, 1), n))
    peaks, avg = torch.full((n,), 123.0), torch.full((n,), 456.0)
    shard.process(local, total_frames, rows, peaks, avg, reduce=False)
    shard.reduce(total_frames, rows, peaks, avg)          # the method the GPU path calls (bench.py, one collective)
    peaks2, avg2 = torch.full((n,), 123.0), torch.full((n,), 456.0)
    shard.process(local, total_frames, rows, peaks2, avg2, reduce=True)
    assert torch.equal(peaks, peaks2) and torch.equal(avg, avg2)
    q.put((total_frames, L, rank, peaks.numpy(), avg.numpy()))


def test_two_rank_reduction_over_gloo():
    """ShardedSpectrum.reduce itself, world size 2: the single-owner branch (one packed all-reduce) and the
    straddling-tail branch (gather + average_rows), against the sequential whole-recording result."""
    world, n = 2, 32
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world * len(CASES))]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(results) == world * len(CASES)
    for total_frames, L, _, peaks, avg in results:
        rng = np.random.default_rng(99)
        rows_all = (rng.standard_normal((total_frames, n)) * 10 - 60).astype(np.float32)
        want_avg = torch.empty(n)
        _RowsPlan.average_rows(torch.from_numpy(rows_all), total_frames - 1, -1, 0, n, total_frames, L, n, want_avg)
        assert np.array_equal(peaks, rows_all.max(axis=0)), (total_frames, L)
        assert np.array_equal(avg, want_avg.numpy()), (total_frames, L)


def test_assemble_tail_orders_rows_newest_first():
    L, n, total, world = 4, 3, 7, 3       # segments 3, 2, 2 frames: the newest 5 rows come from ranks 2, 1, 0
    rows_all = np.arange(total * n, dtype=np.float32).reshape(total, n)
    gathered = []
    for r in range(world):
        first, nloc = shard_frames(total, world, r)
        mine = torch.full((L + 1, n), -9999.0)
        take = min(L + 1, nloc)
        mine[:take] = torch.flip(torch.from_numpy(rows_all[first:first + nloc])[nloc - take:], dims=[0])
        gathered.append(mine)
    tail = assemble_tail(gathered, total, world, L).numpy()
    assert np.array_equal(tail, rows_all[::-1][:L + 1])


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
    interpolation, decimation, taps_per_phase = 1, 25, 10

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


def test_halo_covers_the_delay_lines():
    """one packet of halo is only enough when the packet spans every delay line; at a large
    decimation with small packets the default grows, and an explicit halo that is too short is refused."""
    from rfanalyzer_b200.sharding import delay_line_span, halo_packets_for

    class Big(_FakePlan):      # 20 Msps -> 48 kHz: I/D = 3/1250, 500 taps per phase, 1024-sample packets
        class desc:
            packet_samples, mode, format = 1024, 2, 0
        interpolation, decimation, taps_per_phase = 3, 1250, 500

    span = delay_line_span(Big())
    assert span >= 500 + 26 * 1250 // 3
    assert halo_packets_for(Big()) * 1024 >= span and halo_packets_for(Big()) > 1
    assert halo_packets_for(_FakePlan()) == 2          # 10 + 59 * 25 + 1 samples -> two 1000-sample packets
    with pytest.raises(ValueError):
        ShardedChain(Big(), 0, 2, halo_packets=1)
