"""Edge-case frames and packets through the GPU paths, against the oracle ("What's weak" 1):

* all-zero frames: the reference's `10*log10(sqrt(0))` is -inf (nativedsp.cpp:72-79); rows must be -inf in the same
  bins, peaks stay at FftProcessor's -999999f (max(-999999, -inf)), the box-car average becomes -inf, nothing is NaN;
* full-scale frames (-128 / 127, 0x8000 / 0x7fff, 0 / 255), a single impulse, pure DC, the Nyquist pattern;
* a silent first packet through AM / SSB / CW: the AGC maximum is still zero (Demodulator.kt:285-306, :347-355), so the
  reference emits 0 * (0.75 / 0) = NaN -- the GPU must emit NaN in the same samples and recover with the reference;
* the exponential-averaging option (RFA_AVG_EMA; not in the reference) against its sequential definition.

Tolerances: 0.01 dB where the reference is within 35 dB-units (10*log10 of the AMPLITUDE ratio, the reference's
scale) of the frame's strongest bin; below that a float32 FFT's own rounding noise is the value (pffft built for SSE
and for NEON disagree there too), so only "equally negligible" is asserted, through the linear-power bound."""
import numpy as np
import pytest

from test_spectrum_gpu import DB_TOL, gpu_spectrum

pytestmark = pytest.mark.gpu


def _codes(fmt, n_values):
    """dtype view helpers: (numpy dtype of one component, bytes per sample)."""
    return (np.int8, 2) if fmt == 0 else ((np.uint8, 2) if fmt == 1 else (np.dtype("<i2"), 4))


def _frame(fmt, n, kind):
    dt, _ = _codes(fmt, n)
    lo, hi, zero = {0: (-128, 127, 0), 1: (0, 255, 127), 2: (-32768, 32767, 0)}[fmt]
    x = np.full((n, 2), zero, dtype=np.int64)
    if kind == "zero":
        pass
    elif kind == "full_neg":
        x[:] = lo
    elif kind == "full_pos":
        x[:] = hi
    elif kind == "impulse":
        x[n // 3, 0] = hi
        x[n // 3, 1] = lo
    elif kind == "dc":
        x[:, 0] = hi // 2
        x[:, 1] = lo // 3
    elif kind == "nyquist":
        x[0::2] = hi
        x[1::2] = lo
    elif kind == "one_lsb":
        x[:, 0] = zero + 1
    else:
        raise ValueError(kind)
    return x.astype(dt).reshape(-1).view(np.uint8)


KINDS = ["full_neg", "full_pos", "impulse", "dc", "nyquist", "one_lsb", "zero"]   # silence among the newest L+1 rows


def _compare(rows, ref):
    assert not np.isnan(rows).any() and not np.isnan(ref).any()
    assert np.array_equal(np.isneginf(ref).all(axis=1), np.isneginf(rows).all(axis=1))   # all-zero frames: -inf rows
    for r, g in zip(ref, rows):
        if np.isneginf(r).all():
            assert np.isneginf(g).all()
            continue
        top = r.max()
        assert abs(g.max() - top) < DB_TOL
        strong = r >= top - 35.0
        assert np.abs(g[strong] - r[strong]).max() < DB_TOL
        # everywhere: linear power within 1e-4 relative + 1e-6 of the strongest bin (-inf is power 0)
        lin_g, lin_r = 10.0 ** (g.astype(np.float64) / 5.0), 10.0 ** (r.astype(np.float64) / 5.0)
        assert np.all(np.abs(lin_g - lin_r) <= 1e-4 * lin_r + 1e-6 * lin_r.max())


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("n", [256, 1024, 4096, 8192, 16384, 65536])
@pytest.mark.parametrize("device", [True, False])
def test_edge_frames_vs_oracle(gpu_ctx, oracle, fmt, n, device):
    """Every special frame, surrounded by ordinary ones so that peaks and the average see a mix."""
    if not device and n not in (4096, 65536):
        pytest.skip("host-buffer path: two sizes are enough")
    bps = 2 if fmt < 2 else 4
    normal = oracle.synth_iq(fmt, n * 2)
    parts = [normal[: n * bps]] + [_frame(fmt, n, k) for k in KINDS] + [normal[n * bps:]]
    iq = np.concatenate(parts)
    L = 3
    r, p, a = oracle.spectrum_run(fmt, iq, n, L)
    rows, peaks, avg = gpu_spectrum(gpu_ctx, fmt, iq, n, L=L, device=device)
    _compare(rows, r)
    assert not np.isnan(peaks).any() and not np.isnan(avg).any()
    assert np.array_equal(peaks, rows.max(axis=0))
    assert np.abs(peaks - p).max() < DB_TOL
    # the average is an exact function of the rows (AnalyzerSurface.kt:710-714: newest -> oldest, float32) ...
    s = np.zeros(n, np.float32)
    with np.errstate(invalid="ignore"):
        for k in range(L + 1):
            s = (s + rows[len(rows) - 1 - k]).astype(np.float32)
        assert np.array_equal(avg, (s / np.float32(L + 1)).astype(np.float32))
    # ... and agrees with the oracle's wherever its newest L+1 rows are all well above the rounding floor
    tail = r[-(L + 1):]
    solid = np.all(tail >= tail.max(axis=1, keepdims=True) - 35.0, axis=0) & np.isfinite(a)
    if solid.any():
        assert np.abs(avg[solid] - a[solid]).max() < DB_TOL
    if fmt != 1:                 # a silent frame among the newest rows: the average is -inf in every bin
        assert np.isneginf(avg).all() and np.isneginf(a).all()


@pytest.mark.parametrize("fmt", [0, 2])
@pytest.mark.parametrize("n", [1024, 4096, 32768])
def test_all_zero_recording(gpu_ctx, oracle, fmt, n):
    """Nothing but silence: every row is -inf, the peak hold never leaves -999999f (FftProcessor.kt:236,244:
    max(-999999, -inf)), the average is -inf; no NaN anywhere, also on a second accumulating call."""
    import torch
    import rfanalyzer_b200 as rfa
    frames = 9
    iq = np.concatenate([_frame(fmt, n, "zero")] * frames)
    r, p, a = oracle.spectrum_run(fmt, iq, n, 4)
    assert np.isneginf(r).all() and np.all(p == -999999.0) and np.isneginf(a).all()
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n, avg_len=4)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        rows = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
        peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
        avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        for call in range(2):
            plan.process(d, frames, rows=rows, peaks=peaks, avg=avg, peaks_accumulate=call > 0)
            gpu_ctx.sync()
            assert torch.isneginf(rows).all() and torch.isneginf(avg).all()
            assert torch.all(peaks == -999999.0)
        # a real frame after the silence: peaks become that frame's row, the average stays -inf (it still spans silence)
        live = oracle.synth_iq(fmt, n)
        both = torch.from_numpy(np.concatenate([iq[: 2 * len(live)], live])).cuda()
        rows3 = torch.zeros((3, n), dtype=torch.float32, device="cuda")
        plan.process(both, 3, rows=rows3, peaks=peaks, avg=avg, peaks_accumulate=True)
        gpu_ctx.sync()
        assert torch.equal(peaks, rows3[2]) and torch.isfinite(rows3[2]).all() and torch.isneginf(avg).all()


SILENT = [(2, 10_000_000, 1, 8000, 65536), (2, 10_000_000, 5, 2800, 65536), (2, 10_000_000, 4, 2800, 65536),
          (2, 10_000_000, 6, 300, 65536), (0, 2_000_000, 1, 8000, 131072 // 8)]


@pytest.mark.parametrize("fmt,fs,mode,width,packet", SILENT)
@pytest.mark.parametrize("exact", [True, False])
def test_silent_first_packets_through_am_ssb_cw(gpu_ctx, oracle, fmt, fs, mode, width, packet, exact):
    """Two packets of exact silence, then signal: while the AGC maximum is zero the reference's
    `x * 0.75f / lastMax` (Demodulator.kt:299-302, :350-354) is 0/0 -> NaN; the GPU chain emits NaN in exactly those
    samples and the same finite audio afterwards."""
    import rfanalyzer_b200 as rfa
    from test_dsp_gpu import _chain_input
    npk = 5
    n = packet * npk
    iq, src, chan = _chain_input(oracle, rfa, fmt, fs, mode, n)
    bps = 2 if fmt < 2 else 4
    iq = iq.copy()
    iq[: 2 * packet * bps] = 0
    want = oracle.chain_run(fmt, iq, fs, src, chan, mode, width, packet, volume=1.0)
    assert np.isnan(want).any() and np.isfinite(want).any()      # the scenario is what it claims to be
    plan = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 1.0, rfa.SUM_EXACT if exact else rfa.SUM_FMA)
    audio = np.zeros(plan.max_audio(n), np.float32)
    got = audio[: plan.process(iq, n, audio)]
    assert len(got) == len(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = np.isfinite(want)
    if exact:
        assert np.array_equal(got[ok], want[ok])
    else:
        assert np.abs(got[ok] - want[ok]).max() <= 1e-4 * np.abs(want[ok]).max()


def test_many_packets_in_one_call(gpu_ctx, oracle):
    """more than 65535 packets in one AM call (gridDim.y is capped at 65535)."""
    import rfanalyzer_b200 as rfa
    import torch
    fmt, fs, packet = 1, 2_400_000, 64
    npk = 70_000
    n = packet * npk
    plan = rfa.ChainPlan(gpu_ctx, fmt, fs, 100_000_000, 100_240_000, rfa.MODE_AM, 8000, packet, 1.0, rfa.SUM_FMA)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        iq = torch.empty(n * 2, dtype=torch.uint8, device="cuda")
        rfa.synth_iq(gpu_ctx, fmt, n, iq)
        audio = torch.zeros(plan.max_audio(n), dtype=torch.float32, device="cuda")
        got = plan.process(iq, n, audio)
        gpu_ctx.sync()
    assert got > 0
    a = audio[:got].cpu().numpy()
    assert np.isfinite(a[got // 2:]).all() and np.abs(a[got // 2:]).max() > 0
    # the tail packets (index > 65535) were normalised too: bounded by the AGC's 0.75 / lastMax scaling
    assert np.abs(a[-2000:]).max() <= 1.5


# ---- exponential averaging (RFA_AVG_EMA) -----------------------------------------------------------------
@pytest.mark.parametrize("fmt,n,frames,alpha", [(0, 4096, 64, 0.25), (1, 1024, 300, 0.05), (2, 16384, 9, 1.0),
                                                (0, 65536, 12, 0.5), (0, 4096, 1, 0.3), (0, 256, 4000, 0.02)])
@pytest.mark.parametrize("device", [True, False])
def test_ema_vs_sequential_definition(gpu_ctx, oracle, fmt, n, frames, alpha, device):
    import torch
    import rfanalyzer_b200 as rfa
    iq = oracle.synth_iq(fmt, n * frames)
    r, p, _ = oracle.spectrum_run(fmt, iq, n, 0)
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n, avg_len=8, ema_alpha=alpha)
    if device:
        with torch.cuda.stream(gpu_ctx.torch_stream):
            d = torch.from_numpy(iq).cuda()
            rows = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
            peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
            avg = torch.zeros(n, dtype=torch.float32, device="cuda")
            plan.process(d, frames, rows=rows, peaks=peaks, avg=avg)
            gpu_ctx.sync()
            avg2 = torch.zeros(n, dtype=torch.float32, device="cuda")
            plan.process(d, frames, rows=None, peaks=peaks, avg=avg2)        # no rows kept: the plan's own window
            gpu_ctx.sync()
        rows, avg, avg2 = rows.cpu().numpy(), avg.cpu().numpy(), avg2.cpu().numpy()
    else:
        rows = np.zeros((frames, n), np.float32)
        peaks, avg = np.zeros(n, np.float32), np.zeros(n, np.float32)
        with gpu_ctx.options(chunk_kib=64):      # several chunks: the average is carried from chunk to chunk
            plan.process(iq, frames, rows=rows, peaks=peaks, avg=avg)
        avg2 = avg
    # exactly the sequential recurrence over the GPU's own rows whenever the window covers the call ...
    want_own = oracle.ema_rows(rows, alpha)
    assert np.abs(avg - want_own).max() <= 2e-5
    assert np.abs(avg2 - want_own).max() <= 2e-5
    if alpha >= 0.25 or frames < 100:
        assert np.array_equal(avg, want_own)
    # ... and within the dB tolerance of the oracle's rows
    assert np.abs(avg - oracle.ema_rows(r, alpha)).max() < DB_TOL


def test_ema_carries_across_calls(gpu_ctx, oracle):
    """avg_accumulate continues the recurrence from the caller's vector: two calls equal one (ring and linear)."""
    import torch
    import rfanalyzer_b200 as rfa
    n, alpha, frames = 2048, 0.2, 40
    iq = oracle.synth_iq(1, n * frames)
    plan = rfa.SpectrumPlan(gpu_ctx, 1, n, ema_alpha=alpha)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        rows = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
        one = torch.zeros(n, dtype=torch.float32, device="cuda")
        plan.process(d, frames, rows=rows, avg=one)
        two = torch.zeros(n, dtype=torch.float32, device="cuda")
        k = 13
        ring = torch.full((300, n), -9999.0, dtype=torch.float32, device="cuda")
        plan.process(d[: k * n * 2], k, rows=ring, avg=two, row0=0, row_step=-1, ring_rows=300)
        plan.process(d[k * n * 2:], frames - k, rows=ring, avg=two, row0=(0 - k) % 300, row_step=-1, ring_rows=300,
                     history_rows=k, avg_accumulate=True)
        gpu_ctx.sync()
    assert torch.equal(one, two)
    assert np.array_equal(one.cpu().numpy(), oracle.ema_rows(rows.cpu().numpy(), alpha))
    # the stand-alone entry point
    out = np.zeros(n, np.float32)
    gpu_ctx.ema_rows(rows, 0, 1, 0, n, 0, frames - 1, alpha, False, n, out)
    assert np.array_equal(out, one.cpu().numpy())
