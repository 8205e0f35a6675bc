"""Scan detectors: the C ABI's host arithmetic against the oracle restatement of
ui/MainViewModel.kt on the CPU, and the window reductions on the GPU against the same oracle.

The reference has no tests for these functions ("parity unpinned by reference tests"); the oracle follows the
cited lines with float32 / double arithmetic exactly where Kotlin has Float / Double."""
import ctypes as C

import numpy as np
import pytest

from oracle import detectors as D


@pytest.fixture(scope="module")
def lib():
    from rfanalyzer_b200 import _lib
    return _lib.load()


def test_bin_and_half_width_match_the_jvm_arithmetic(lib):
    rng = np.random.default_rng(1)
    for _ in range(3000):
        n = int(2 ** rng.integers(4, 17))
        fs = int(rng.integers(100_000, 61_440_000))
        center = int(rng.integers(1_000_000, 6_000_000_000))
        freq = center + int(rng.integers(-fs, fs))
        res = D.resolution(fs, n)
        assert lib.rfa_detect_bin(center, fs, n, freq) == D.bin_index(freq, center - D.jdiv(fs, 2), res)
        for hz, mn in ((100000, 5), (12500, 3)):
            assert lib.rfa_detect_half_width(fs, n, hz, mn) == max(D.to_int(D.F(D.F(hz) / res)), mn)
    # far outside the row: toInt() saturates instead of wrapping
    assert lib.rfa_detect_bin(0, 1, 65536, 10 ** 15) == 2147483647
    assert lib.rfa_detect_bin(10 ** 15, 1, 65536, 0) == -2147483648


def test_window_at_clamps_like_coerce(lib):
    b, s, e = C.c_int(), C.c_int(), C.c_int()
    n, fs, center = 4096, 2_500_000, 600_000_000
    assert lib.rfa_detect_window_at(center, fs, n, center - fs // 2 + 1000, 100000, 5, C.byref(b), C.byref(s), C.byref(e)) == 1
    assert (b.value, s.value) == (1, 0) and e.value == 1 + 163
    assert lib.rfa_detect_window_at(center, fs, n, center + fs, 100000, 5, C.byref(b), C.byref(s), C.byref(e)) == 0
    assert lib.rfa_detect_window_at(center, fs, n, center, -1, 2, C.byref(b), C.byref(s), C.byref(e)) == 1
    assert (b.value, s.value, e.value) == (2048, 2046, 2050)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_decide(lib, mode):
    rng = np.random.default_rng(mode)
    for _ in range(2000):
        peak, avg, thr, nf, mg = (float(np.float32(v)) for v in rng.uniform(-120, 0, 5))
        assert lib.rfa_detect_decide(peak, avg, thr, nf, mg, mode) == int(D.decide(np.float32(peak), np.float32(avg), thr, nf, mg, mode))
    assert lib.rfa_detect_decide(0.0, 0.0, 0.0, 0.0, 0.0, 7) == -1


def test_scan_grid_matches_the_reference_loop(lib):
    from rfanalyzer_b200 import detect
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = int(2 ** rng.integers(8, 17))
        fs = int(rng.integers(1_000_000, 20_000_000))
        center = int(rng.integers(50_000_000, 2_000_000_000))
        usable = int(fs * rng.uniform(0.5, 1.0))
        step = int(rng.choice([5000, 12500, 25000, 100000, 200000]))
        lo, hi = center - int(rng.integers(0, fs)), center + int(rng.integers(0, fs))
        freqs, wins, count = detect.scan_grid(center, fs, usable, step, lo, hi, n, row=7)
        row = np.zeros(n, np.float32)   # every grid point "detected": threshold below the row
        ref = D.detect_signals_in_fft(row, center, fs, usable, step, -10.0, 0, -100.0, 0.0, lo, hi)
        assert count == len(ref)
        assert [int(f) for f in freqs] == [r[0] for r in ref]
        res = D.resolution(fs, n)
        for i in range(count):
            b = D.bin_index(int(freqs[i]), center - D.jdiv(fs, 2), res)
            assert (wins[i].row, wins[i].start, wins[i].end) == (7, max(0, b - 2), min(n - 1, b + 2))


def test_group_signals(lib):
    from rfanalyzer_b200 import detect
    rng = np.random.default_rng(5)
    assert detect.groupSignals([], 25000, 2) == []
    for _ in range(200):
        step, gap = int(rng.choice([12500, 25000, 100000])), int(rng.integers(1, 4))
        k = int(rng.integers(1, 40))
        freqs = 100_000_000 + step * rng.choice(200, size=k, replace=False)
        sig = [detect.DiscoveredSignal(int(f), float(np.float32(rng.uniform(-90, -10))), float(np.float32(rng.uniform(-110, -40))))
               for f in freqs]
        got = detect.groupSignals(sig, step, gap)
        ref = D.group_signals([(s.frequency, np.float32(s.peakStrength), np.float32(s.averageStrength), 0, False) for s in sig], step, gap)
        assert len(got) == len(ref)
        for g, r in zip(got, ref):
            assert (g.frequency, g.bandwidth, g.isGrouped) == (r[0], r[3], r[4])
            assert np.float32(g.peakStrength) == np.float32(r[1]) and np.float32(g.averageStrength) == np.float32(r[2])


def test_squelch():
    from rfanalyzer_b200 import detect
    assert detect.squelchSatisfied(-30.0, -40.0, True) and not detect.squelchSatisfied(-50.0, -40.0, True)
    assert detect.squelchSatisfied(-999.0, -40.0, False)


# ------------------------------------------------------------------------------------------ GPU
def _ring(gpu_ctx, oracle, n, ring, fmt=0):
    """A device ring filled by the spectrum path itself + its host copy."""
    import torch
    import rfanalyzer_b200 as rfa
    iq = oracle.synth_iq(fmt, n * ring)
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        rows = torch.zeros((ring, n), dtype=torch.float32, device="cuda")
        plan.process(torch.from_numpy(iq).cuda(), ring, rows=rows)
        gpu_ctx.sync()
    return rows, rows.cpu().numpy()


def _same(a, b):
    a, b = np.float32(a), np.float32(b)
    return (np.isnan(a) and np.isnan(b)) or a == b


@pytest.mark.gpu
@pytest.mark.parametrize("n", [256, 4096, 65536])
def test_window_reductions_vs_oracle(gpu_ctx, oracle, n):
    """peak bit-exact (a max), average equal to the sequential double sum after the float rounding."""
    rows, host = _ring(gpu_ctx, oracle, n, 6)
    rng = np.random.default_rng(n)
    wins = [(r, 0, n - 1) for r in range(6)]
    for _ in range(300):
        a = int(rng.integers(0, n))
        wins.append((int(rng.integers(0, 6)), a, min(n - 1, a + int(rng.integers(0, 400)))))
    wins += [(2, -5, 3), (3, n - 2, n + 50), (1, 17, 17)]       # clamped and single-bin windows
    peak, avg = gpu_ctx.detect_windows(rows, rows.stride(0), n, wins)
    for (r, a, b), p, m in zip(wins, peak, avg):
        w = host[r, max(a, 0):min(b, n - 1) + 1]
        assert _same(p, D.max_or_null(w))
        assert _same(m, np.float32(D.average(w)))


@pytest.mark.gpu
def test_window_reductions_special_values_and_device_outputs(gpu_ctx):
    import torch
    n = 1024
    host = np.random.default_rng(0).uniform(-120, -20, (3, n)).astype(np.float32)
    host[0, 100:110] = -np.inf      # an all-zero frame's bins: log10f(0)
    host[1, 500] = np.nan
    with torch.cuda.stream(gpu_ctx.torch_stream):
        rows = torch.from_numpy(host).cuda()
        wins = [(0, 90, 120), (0, 100, 109), (1, 490, 510), (1, 0, 499), (2, 0, n - 1)]
        peak = torch.zeros(len(wins), dtype=torch.float32, device="cuda")
        avg = torch.zeros(len(wins), dtype=torch.float32, device="cuda")
        gpu_ctx.detect_windows(rows, n, n, wins, peak, avg)
        gpu_ctx.sync()
    for (r, a, b), p, m in zip(wins, peak.cpu().numpy(), avg.cpu().numpy()):
        w = host[r, a:b + 1]
        assert _same(p, D.max_or_null(w)) and _same(m, np.float32(D.average(w)))


@pytest.mark.gpu
def test_reference_named_detectors_vs_oracle(gpu_ctx, oracle):
    import rfanalyzer_b200 as rfa
    from rfanalyzer_b200 import detect
    from types import SimpleNamespace as NS
    n, ring, fs, center = 4096, 8, 20_000_000, 100_000_000
    rows, host = _ring(gpu_ctx, oracle, n, ring)
    d = rfa.FftProcessorData()
    d.waterfallBuffer, d.readIndex = rows, 5
    row = host[5]
    assert _same(detect.getAverageSignalLevel(gpu_ctx, d), D.get_average_signal_level(row))
    nf = float(D.get_average_signal_level(row))
    for mode in (0, 1, 2):
        got, ref = detect.detectSignal(gpu_ctx, d, -60.0, mode, nf, 10.0), D.detect_signal(row, -60.0, mode, nf, 10.0)
        assert (got is None) == (ref is None)
        if got:
            assert _same(got[0], ref[0]) and _same(got[1], ref[1])
        got = detect.detectSignalsInFFT(gpu_ctx, d, center, fs, 16_000_000, 25_000, -200.0, mode, nf, 8.0, 90_000_000, 109_000_000)
        ref = D.detect_signals_in_fft(row, center, fs, 16_000_000, 25_000, -200.0, mode, nf, 8.0, 90_000_000, 109_000_000)
        assert len(got) == len(ref) and len(ref) > 0      # the synthetic tones stand well above the noise floor
        for g, r in zip(got, ref):
            assert g.frequency == r[0] and _same(g.peakStrength, r[1]) and _same(g.averageStrength, r[2])
        grouped = detect.groupSignals(got, 25_000, 2)
        gref = D.group_signals(ref, 25_000, 2)
        assert [(g.frequency, g.bandwidth) for g in grouped] == [(r[0], r[3]) for r in gref]
    # a batch of rows in one launch equals the row-by-row answers
    batch = detect.detectSignalsInFFT_rows(gpu_ctx, rows, list(range(ring)), center, fs, 16_000_000, 25_000, -200.0, 0, nf, 8.0,
                                           90_000_000, 109_000_000)
    for r in range(ring):
        ref = D.detect_signals_in_fft(host[r], center, fs, 16_000_000, 25_000, -200.0, 0, nf, 8.0, 90_000_000, 109_000_000)
        assert [(s.frequency, np.float32(s.peakStrength)) for s in batch[r]] == [(t[0], np.float32(t[1])) for t in ref]
    channels = [NS(id=i, frequency=center - fs // 2 + int(f)) for i, f in enumerate(np.linspace(-1e6, 21e6, 40))]
    got = detect.detectIEMChannelsInFFT(gpu_ctx, d, channels, center, fs, nf + 10.0)
    ref = D.detect_iem_channels(row, [(c.id, c.frequency) for c in channels], center, fs, nf + 10.0)
    assert [g.channelId for g in got] == [r[0] for r in ref] and len(ref) > 0
    for g, r in zip(got, ref):
        assert _same(g.peakStrength, r[1]) and _same(g.averageStrength, r[2])
    assert _same(detect.detectAirCommSignal(gpu_ctx, d, fs), D.detect_aircomm_signal(row, fs))
    for target in (center + 2_468_000, center - 9_999_999, center + 30_000_000):
        got, ref = detect.detectAirCommSignalAtFrequency(gpu_ctx, d, target, center, fs), D.detect_aircomm_signal_at_frequency(row, target, center, fs)
        assert (got is None) == (ref is None) and (got is None or _same(got, ref))
    d.readIndex = ring      # the reference's guard: readIndex outside the ring -> no answer
    assert detect.getAverageSignalLevel(gpu_ctx, d) is None and detect.detectSignalsInFFT(gpu_ctx, d, center, fs, fs, 25000, -200.0, 0, nf, 8.0, 0, 10 ** 10) == []
