"""The fused spectrum path on the GPU, through the C ABI, against the oracle, the compiled
reference's golden rows, and size-independent properties at the benchmark's full size.

Tolerances are BASELINE.json's: 0.01 dB on log spectra; 1e-4 relative on linear power with
an absolute floor of 1e-6 of the frame's strongest bin (bins ~100 dB under the peak are float32
rounding noise in the reference itself: pffft-SSE and pffft-NEON do not agree there either)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DB_TOL = 0.01


@pytest.fixture(autouse=True)
def _default_knobs(request):
    """Tests may turn the context's knobs (rfa_ctx_set_option); every test starts from the defaults."""
    yield
    if "gpu_ctx" in request.fixturenames:
        ctx = request.getfixturevalue("gpu_ctx")
        for k, v in (("fs_batch_kib", 128 << 10), ("fs_tma", 1), ("fs_ztma", 1), ("staged", 1), ("pdl", 1), ("cluster", 0)):
            ctx.set_option(k, v)


def lin_ok(rows, ref):
    lin, lin_ref = 10.0 ** (rows.astype(np.float64) / 5.0), 10.0 ** (ref.astype(np.float64) / 5.0)
    floor = 1e-6 * lin_ref.max(axis=-1, keepdims=True)
    return np.all(np.abs(lin - lin_ref) <= 1e-4 * lin_ref + floor)


def gpu_spectrum(ctx, fmt, iq, n, L=0, device=True, window=0, peak_hold=True):
    import torch
    import rfanalyzer_b200 as rfa
    frames = len(iq) // (n * rfa.BYTES_PER_SAMPLE[fmt])
    plan = rfa.SpectrumPlan(ctx, fmt, n, window=window, avg_len=L, peak_hold=peak_hold)
    if device:
        with torch.cuda.stream(ctx.torch_stream):
            d = torch.from_numpy(iq).cuda()
            rows = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
            peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
            avg = torch.zeros(n, dtype=torch.float32, device="cuda")
            plan.process(d, frames, rows=rows, peaks=peaks, avg=avg)
            ctx.sync()
        return rows.cpu().numpy(), peaks.cpu().numpy(), avg.cpu().numpy()
    rows = np.zeros((frames, n), np.float32)
    peaks, avg = np.zeros(n, np.float32), np.zeros(n, np.float32)
    plan.process(iq, frames, rows=rows, peaks=peaks, avg=avg)
    return rows, peaks, avg


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("n", [16, 64, 256, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
@pytest.mark.parametrize("device", [True, False])
def test_rows_peaks_avg_vs_oracle(gpu_ctx, oracle, fmt, n, device):
    frames = 37 if n <= 4096 else 5
    iq = oracle.synth_iq(fmt, n * frames)
    r, p, a = oracle.spectrum_run(fmt, iq, n, 3)
    rows, peaks, avg = gpu_spectrum(gpu_ctx, fmt, iq, n, L=3, device=device)
    assert np.abs(rows - r).max() < DB_TOL
    assert np.abs(peaks - p).max() < DB_TOL
    assert np.abs(avg - a).max() < DB_TOL
    assert lin_ok(rows, r)
    # bin indexing is exact: the strongest bin of every frame is the same bin
    assert np.array_equal(rows.argmax(axis=1), r.argmax(axis=1))


def test_against_compiled_reference_golden(gpu_ctx, oracle):
    """Rows produced by the reference's own pffft.c + nativedsp.cpp (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "spectrum_ref.npz"))
    for fmt, name in ((0, "s8"), (1, "u8"), (2, "s16")):
        for n in (1024, 4096, 65536):
            frames = 2 if n <= 4096 else 1
            iq = oracle.synth_iq(fmt, n * frames, first=12345)
            rows, peaks, avg = gpu_spectrum(gpu_ctx, fmt, iq, n, L=1)
            assert np.abs(rows - g[f"{name}_{n}_rows"]).max() < DB_TOL
            assert np.abs(peaks - g[f"{name}_{n}_peaks"]).max() < DB_TOL
            assert np.abs(avg - g[f"{name}_{n}_avg"]).max() < DB_TOL
            assert lin_ok(rows, g[f"{name}_{n}_rows"])


def test_average_and_peaks_are_exact_functions_of_the_rows(gpu_ctx, oracle):
    """peaks = element-wise max of the rows (FftProcessor.kt:244), avg = float32 sum of the newest
    L+1 rows, newest first, / (L+1) (AnalyzerSurface.kt:710-714): bit-exact given the GPU's rows."""
    n, frames, L = 4096, 50, 8
    iq = oracle.synth_iq(0, n * frames)
    rows, peaks, avg = gpu_spectrum(gpu_ctx, 0, iq, n, L=L)
    assert np.array_equal(peaks, rows.max(axis=0))
    s = np.zeros(n, np.float32)
    for k in range(L + 1):
        s = (s + rows[frames - 1 - k]).astype(np.float32)
    assert np.array_equal(avg, (s / np.float32(L + 1)).astype(np.float32))


def test_fewer_frames_than_average_length(gpu_ctx, oracle):
    """Rows that were never written count as -9999f (FftProcessor.kt:181)."""
    n, frames, L = 1024, 2, 4
    iq = oracle.synth_iq(0, n * frames)
    rows, _, avg = gpu_spectrum(gpu_ctx, 0, iq, n, L=L)
    s = np.zeros(n, np.float32)
    for v in (rows[1], rows[0], np.full(n, -9999.0, np.float32), np.full(n, -9999.0, np.float32),
              np.full(n, -9999.0, np.float32)):
        s = (s + v).astype(np.float32)
    assert np.array_equal(avg, (s / np.float32(5)).astype(np.float32))


def test_reference_ring_layout_and_accumulating_peaks(gpu_ctx, oracle):
    """FftProcessorData semantics across calls: rows run backwards through a 300-row ring
    (FftProcessor.kt:224-229), peaks keep accumulating, the average sees earlier calls' rows."""
    import torch
    import rfanalyzer_b200 as rfa
    n, L, ring = 1024, 5, 300
    L_o = oracle.lib()
    proc = L_o.orc_fftproc_new(ring, 1)
    plan = rfa.SpectrumPlan(gpu_ctx, 0, n, avg_len=L)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d_ring = torch.full((ring, n), -9999.0, dtype=torch.float32, device="cuda")
        d_peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
        d_avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        write_index, history, first = 0, 0, 0
        for call, frames in enumerate((3, 1, 310, 7)):
            iq = oracle.synth_iq(0, n * frames, first=first)
            first += n * frames
            r, _, _ = oracle.spectrum_run(0, iq, n, 0)
            for k in range(frames):
                L_o.orc_fftproc_push(proc, np.ascontiguousarray(r[k]), n, 100_000_000, 20_000_000)
            plan.process(torch.from_numpy(iq).cuda(), frames, rows=d_ring, peaks=d_peaks, avg=d_avg,
                         row0=write_index, row_step=-1, ring_rows=ring, history_rows=history,
                         peaks_accumulate=call > 0)
            gpu_ctx.sync()
            write_index = (write_index - frames) % ring
            history = min(ring, history + frames)
            assert write_index == L_o.orc_fftproc_write_index(proc)
            ring_ref = np.stack([np.ctypeslib.as_array(L_o.orc_fftproc_row(proc, i), shape=(n,)) for i in range(ring)])
            assert np.abs(d_ring.cpu().numpy() - ring_ref).max() < DB_TOL
            peaks_ref = np.ctypeslib.as_array(L_o.orc_fftproc_peaks(proc), shape=(n,))
            assert np.abs(d_peaks.cpu().numpy() - peaks_ref).max() < DB_TOL
            avg_ref = np.empty(n, np.float32)
            L_o.orc_time_average(proc, L, avg_ref)
            assert np.abs(d_avg.cpu().numpy() - avg_ref).max() < DB_TOL
    L_o.orc_fftproc_free(proc)


def test_no_rows_mode_and_peak_hold_off(gpu_ctx, oracle):
    import torch
    import rfanalyzer_b200 as rfa
    n, frames, L = 2048, 40, 3
    iq = oracle.synth_iq(1, n * frames)
    r, p, a = oracle.spectrum_run(1, iq, n, L)
    plan = rfa.SpectrumPlan(gpu_ctx, 1, n, avg_len=L)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
        avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        plan.process(d, frames, rows=None, peaks=peaks, avg=avg)
        gpu_ctx.sync()
    assert np.abs(peaks.cpu().numpy() - p).max() < DB_TOL and np.abs(avg.cpu().numpy() - a).max() < DB_TOL
    off = rfa.SpectrumPlan(gpu_ctx, 1, n, avg_len=L, peak_hold=False)
    hp = np.full(n, 123.0, np.float32)
    off.process(iq, frames, rows=np.zeros((frames, n), np.float32), peaks=hp)
    assert np.all(hp == 123.0)  # FftProcessor.kt:246-248: peaks untouched (null) when peak hold is off


def test_hann_window_variant(gpu_ctx, oracle):
    """north_star's Hann variant: same kernel, different table; checked against a float64 FFT."""
    n, frames = 4096, 3
    iq = oracle.synth_iq(0, n * frames)
    rows, _, _ = gpu_spectrum(gpu_ctx, 0, iq, n, window=1)
    x = iq.view(np.int8).astype(np.float64).reshape(frames, n, 2) / 128.0
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / (n - 1))).astype(np.float32).astype(np.float64)
    spec = np.fft.fftshift(np.fft.fft((x[..., 0] + 1j * x[..., 1]) * w, axis=1), axes=1)
    ref = 10 * np.log10(np.abs(spec) / n)
    assert np.abs(rows - ref).max() < DB_TOL


@pytest.mark.parametrize("n", [16, 512, 4096, 16384, 65536])
def test_native_dsp_entry_points(gpu_ctx, oracle, n):
    """performFFT / performFFTAndLogMag / performWindowedFftAndReturnMag (nativedsp.cpp:19-81,
    NativeDsp.kt:43-62), host buffers like the JNI arrays."""
    rng = np.random.default_rng(n)
    batch = 3
    x = rng.standard_normal((batch, 2 * n)).astype(np.float32)
    out = np.empty_like(x)
    gpu_ctx.fft_c2c(x, out, n, batch)
    ref = np.fft.fft(x[:, 0::2].astype(np.float64) + 1j * x[:, 1::2].astype(np.float64), axis=1)
    got = out[:, 0::2] + 1j * out[:, 1::2]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-6
    mag = np.empty((batch, n), np.float32)
    gpu_ctx.fft_logmag(x, mag, n, batch)
    L = oracle.lib()
    for b in range(batch):
        m = np.empty(n, np.float32)
        L.orc_fft_logmag(np.ascontiguousarray(x[b]), m, n)
        assert np.abs(mag[b] - m).max() < DB_TOL
    re = np.ascontiguousarray(x[:, :n])
    im = np.ascontiguousarray(x[:, n:])
    gpu_ctx.windowed_fft_logmag(re, im, mag, n, batch)
    for b in range(batch):
        m = np.empty(n, np.float32)
        assert L.orc_windowed_fft_logmag(np.ascontiguousarray(re[b]), np.ascontiguousarray(im[b]), n, n, n, m) == 1
        assert np.abs(mag[b] - m).max() < DB_TOL


def test_unsupported_sizes_fail_loudly(gpu_ctx):
    import rfanalyzer_b200 as rfa
    for bad in (0, 8, 1000, 131072):
        with pytest.raises(rfa.RfaError):
            rfa.SpectrumPlan(gpu_ctx, 0, bad)
    with pytest.raises(rfa.RfaError):
        rfa.SpectrumPlan(gpu_ctx, 0, 4096, avg_len=31)


def test_full_size_properties(gpu_ctx, oracle):
    """BASELINE config 1 at full size (int8, 2^24 samples, N=4096, L=8, peak hold), checked through
    properties that need no full-size oracle run:
      * Parseval per frame: sum |X|^2 = N * sum |x*w|^2          (from the dB rows, 1e-4)
      * sampled frames against the oracle                          (0.01 dB)
      * peaks = max over rows, avg = mean of the newest 9 rows      (bit exact)
      * frames are independent: a second pass over a sub-range reproduces those rows bit for bit."""
    import torch
    import rfanalyzer_b200 as rfa
    n, frames, L = 4096, 4096, 8
    iq = oracle.synth_iq(0, n * frames)
    plan = rfa.SpectrumPlan(gpu_ctx, 0, n, avg_len=L)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        rows = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
        peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
        avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        plan.process(d, frames, rows=rows, peaks=peaks, avg=avg)
        sub = torch.zeros((100, n), dtype=torch.float32, device="cuda")
        plan.process(d[1000 * n * 2:], 100, rows=sub)
        gpu_ctx.sync()
        assert torch.equal(sub, rows[1000:1100])
        assert torch.equal(peaks, rows.max(dim=0).values)
        s = torch.zeros(n, dtype=torch.float32, device="cuda")
        for k in range(L + 1):
            s = s + rows[frames - 1 - k]
        # (torch divides by a scalar through a reciprocal multiply; the reference divides)
        assert np.array_equal(avg.cpu().numpy(), (s.cpu().numpy() / np.float32(L + 1)).astype(np.float32))
        power = (10.0 ** (rows.double() / 5.0)).sum(dim=1) * n * n          # sum |X|^2
    w = oracle.nativedsp_window(n).astype(np.float64)
    x = iq.view(np.int8).astype(np.float64).reshape(frames, n, 2) / 128.0
    energy = ((x[..., 0] ** 2 + x[..., 1] ** 2) * w ** 2).sum(axis=1) * n
    assert np.abs(power.cpu().numpy() / energy - 1.0).max() < 1e-4
    for f in (0, 1, 777, 2048, 4095):
        r, _, _ = oracle.spectrum_run(0, iq[f * n * 2:(f + 1) * n * 2], n, 0)
        assert np.abs(rows[f].cpu().numpy() - r[0]).max() < DB_TOL


@pytest.mark.parametrize("fmt,n,frames,first,count,chunk", [(0, 4096, 300, 0, -1, 64), (1, 1024, 1000, 17, 900, 0),
                                                            (2, 2048, 123, 3, -1, 50), (0, 4096, 5, 0, -1, 2)])
def test_process_file_matches_the_in_memory_pass(gpu_ctx, oracle, tmp_path, fmt, n, frames, first, count, chunk):
    """rfa_spectrum_process_file: reader thread + pinned double buffers; a trailing
    partial frame in the file is ignored; rows / peaks / avg are those of the in-memory call."""
    import rfanalyzer_b200 as rfa
    L = 3
    iq = oracle.synth_iq(fmt, n * frames + 7)                    # 7 samples of a partial frame at the end
    path = os.path.join(tmp_path, "20250111-143022_t_%s_100MHz_6MSps.iq" % ["HACKRF", "RTLSDR", "AIRSPY"][fmt])
    iq.tofile(path)
    total = frames - first if count < 0 else count
    bps = rfa.BYTES_PER_SAMPLE[fmt]
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n, avg_len=L)
    want_rows = np.zeros((total, n), np.float32)
    want_peaks, want_avg = np.zeros(n, np.float32), np.zeros(n, np.float32)
    seg = iq[first * n * bps:(first + total) * n * bps]
    plan.process(seg, total, rows=want_rows, peaks=want_peaks, avg=want_avg)
    rows = np.zeros((total, n), np.float32)
    peaks, avg = np.zeros(n, np.float32), np.zeros(n, np.float32)
    assert plan.process_file(path, first, count, rows=rows, peaks=peaks, avg=avg, chunk_frames=chunk) == total
    assert np.array_equal(rows, want_rows) and np.array_equal(peaks, want_peaks) and np.array_equal(avg, want_avg)
    assert parse_name(gpu_ctx, path)[0] == [0, 1, 2][fmt]


def parse_name(ctx, path):
    import rfanalyzer_b200 as rfa
    return rfa.parse_recording_name(ctx, os.path.basename(path))


# ---- four-step path (fourstep_kernel.cuh), N = 32768 / 65536 ----
@pytest.mark.parametrize("fmt,n,frames,batch_kib,cluster", [(0, 65536, 7, 1024, 0), (2, 32768, 9, 512, 0), (1, 65536, 3, 0, 0),
                                                            (0, 65536, 7, 0, 1), (2, 32768, 9, 0, 1), (1, 32768, 3, 0, 1),
                                                            (2, 65536, 5, 0, 1)])
def test_fourstep_batches_ring_and_history(gpu_ctx, oracle, monkeypatch, fmt, n, frames, batch_kib, cluster):
    """Several batches through a small intermediate buffer (knob "fs_batch_kib"), rows into a
    backwards ring with history, accumulating peaks, average over ring rows -- through the two-kernel path
    (knob "cluster" = 0, the default) and through the cluster path (fourstep_cluster.cuh, "cluster" = 1)."""
    import torch
    import rfanalyzer_b200 as rfa
    L, ring = 4, 12
    gpu_ctx.set_option("cluster", cluster)  # restored by the fixture
    if batch_kib:   # several batches through a small intermediate buffer (context knob, restored by the fixture below)
        gpu_ctx.set_option("fs_batch_kib", batch_kib)
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n, avg_len=L)
    L_o = oracle.lib()
    proc = L_o.orc_fftproc_new(ring, 1)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d_ring = torch.full((ring, n), -9999.0, dtype=torch.float32, device="cuda")
        d_peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
        d_avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        write_index, history, first = 0, 0, 0
        for call, fr in enumerate((frames, 2, 14)):
            iq = oracle.synth_iq(fmt, n * fr, first=first)
            first += n * fr
            r, _, _ = oracle.spectrum_run(fmt, iq, n, 0)
            for k in range(fr):
                L_o.orc_fftproc_push(proc, np.ascontiguousarray(r[k]), n, 100_000_000, 20_000_000)
            plan.process(torch.from_numpy(iq).cuda(), fr, rows=d_ring, peaks=d_peaks, avg=d_avg, row0=write_index,
                         row_step=-1, ring_rows=ring, history_rows=history, peaks_accumulate=call > 0)
            gpu_ctx.sync()
            write_index = (write_index - fr) % ring
            history = min(ring, history + fr)
            ring_ref = np.stack([np.ctypeslib.as_array(L_o.orc_fftproc_row(proc, i), shape=(n,)) for i in range(ring)])
            assert np.abs(d_ring.cpu().numpy() - ring_ref).max() < DB_TOL
            peaks_ref = np.ctypeslib.as_array(L_o.orc_fftproc_peaks(proc), shape=(n,))
            assert np.abs(d_peaks.cpu().numpy() - peaks_ref).max() < DB_TOL
            avg_ref = np.empty(n, np.float32)
            L_o.orc_time_average(proc, L, avg_ref)
            assert np.abs(d_avg.cpu().numpy() - avg_ref).max() < DB_TOL
    L_o.orc_fftproc_free(proc)


@pytest.mark.parametrize("fmt,n", [(0, 65536), (1, 32768), (2, 65536), (2, 32768)])
def test_fourstep_tensor_map_staging_equals_per_thread_loads(gpu_ctx, oracle, monkeypatch, fmt, n):
    """The column kernel reads its raw IQ either through a 2-D tensor-map box per frame (default, 16-byte aligned
    input) or with per-thread loads (knob "fs_tma" = 0, and any misaligned input): same arithmetic, identical rows."""
    import torch
    import rfanalyzer_b200 as rfa
    frames = 5
    iq = oracle.synth_iq(fmt, n * frames)
    gpu_ctx.set_option("cluster", 0)  # the two-kernel path (restored by the fixture)
    rows_t, peaks_t, _ = gpu_spectrum(gpu_ctx, fmt, iq, n, L=2)
    with gpu_ctx.options(fs_tma=0):
        rows_l, peaks_l, _ = gpu_spectrum(gpu_ctx, fmt, iq, n, L=2)
    assert np.array_equal(rows_t, rows_l) and np.array_equal(peaks_t, peaks_l)
    with gpu_ctx.options(fs_ztma=0):   # tensor-map loads, per-thread stores of the intermediate
        rows_s, peaks_s, _ = gpu_spectrum(gpu_ctx, fmt, iq, n, L=2)
    assert np.array_equal(rows_t, rows_s) and np.array_equal(peaks_t, peaks_s)
    # an input that starts 4 bytes into an allocation is not 16-byte aligned: the launcher must fall back by itself
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        buf = torch.zeros(len(iq.view(np.uint8)) + 4, dtype=torch.uint8, device="cuda")
        buf[4:] = torch.from_numpy(iq.view(np.uint8)).cuda()
        rows_m = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
        plan.process(buf[4:], frames, rows=rows_m)
        gpu_ctx.sync()
    assert np.array_equal(rows_m.cpu().numpy(), rows_t)


@pytest.mark.parametrize("fmt,n", [(0, 65536), (1, 32768), (2, 65536), (2, 32768), (0, 32768)])
def test_cluster_path_equals_the_two_kernel_path(gpu_ctx, oracle, fmt, n):
    """N >= 32768 on thread-block clusters (the intermediate in distributed shared memory, ONE launch) against the
    two-kernel path (the intermediate in HBM): the same butterflies in the same order, so identical bits.  More frames
    than clusters, so that every cluster walks several frames and the split cluster barrier is exercised; an input
    that is not 16-byte aligned must fall back to the two-kernel path by itself."""
    import torch
    import rfanalyzer_b200 as rfa
    frames = 2 * 37 + 5 if n == 65536 else 2 * 74 + 3
    iq = oracle.synth_iq(fmt, n * frames)
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n)
    out = {}
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        for cluster in (1, 0):
            gpu_ctx.set_option("cluster", cluster)
            rows = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
            peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
            l0 = gpu_ctx.launch_count
            plan.process(d, frames, rows=rows)
            launches = gpu_ctx.launch_count - l0
            plan.process(d, frames, rows=None, peaks=peaks)
            gpu_ctx.sync()
            out[cluster] = (rows.cpu().numpy(), peaks.cpu().numpy(), launches)
        assert out[1][2] == 1 and out[0][2] == 2      # the cluster path really ran: one launch instead of two
        assert np.array_equal(out[1][0], out[0][0]) and np.array_equal(out[1][1], out[0][1])
        r, p, _ = oracle.spectrum_run(fmt, iq[: 3 * n * rfa.BYTES_PER_SAMPLE[fmt]], n, 0)
        assert np.abs(out[1][0][:3] - r).max() < DB_TOL and lin_ok(out[1][0][:3], r)
        assert np.array_equal(out[1][1], out[1][0].max(axis=0))
        gpu_ctx.set_option("cluster", 1)  # (restored by the fixture)
        buf = torch.zeros(len(iq.view(np.uint8)) + 4, dtype=torch.uint8, device="cuda")
        buf[4:] = torch.from_numpy(iq.view(np.uint8)).cuda()
        rows_m = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
        l0 = gpu_ctx.launch_count
        plan.process(buf[4:], frames, rows=rows_m)
        gpu_ctx.sync()
        assert gpu_ctx.launch_count - l0 == 2
    assert np.array_equal(rows_m.cpu().numpy(), out[1][0])


@pytest.mark.parametrize("cluster", [1, 0])
@pytest.mark.parametrize("n", [32768, 65536])
def test_fourstep_no_rows_peak_only_and_host_buffers(gpu_ctx, oracle, n, cluster):
    """The four-step path behind every output mode of rfa_spectrum_process: no rows (only the newest L+1 rows are
    kept for the average), peaks only, average only, and host buffers (chunked H2D -> kernels -> D2H pipeline)."""
    import torch
    import rfanalyzer_b200 as rfa
    frames, L = 11, 3
    gpu_ctx.set_option("cluster", cluster)  # restored by the fixture
    iq = oracle.synth_iq(0, n * frames)
    r, p, a = oracle.spectrum_run(0, iq, n, L)
    plan = rfa.SpectrumPlan(gpu_ctx, 0, n, avg_len=L)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
        avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        plan.process(d, frames, rows=None, peaks=peaks, avg=avg)
        gpu_ctx.sync()
        assert np.abs(peaks.cpu().numpy() - p).max() < DB_TOL and np.abs(avg.cpu().numpy() - a).max() < DB_TOL
        peaks.zero_()
        plan.process(d, frames, rows=None, peaks=peaks)
        gpu_ctx.sync()
        assert np.abs(peaks.cpu().numpy() - p).max() < DB_TOL
        avg.zero_()
        plan.process(d, frames, rows=None, avg=avg)
        gpu_ctx.sync()
        assert np.abs(avg.cpu().numpy() - a).max() < DB_TOL
    h_rows, h_peaks, h_avg = np.zeros((frames, n), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    plan.process(iq, frames, rows=h_rows, peaks=h_peaks, avg=h_avg)
    assert np.abs(h_rows - r).max() < DB_TOL and np.abs(h_peaks - p).max() < DB_TOL and np.abs(h_avg - a).max() < DB_TOL


@pytest.mark.parametrize("fmt,n,ring", [(0, 256, 300), (1, 1024, 300), (0, 4096, 300), (2, 8192, 40), (0, 32768, 24)])
def test_back_to_back_calls_without_host_synchronisation(gpu_ctx, oracle, fmt, n, ring):
    """Consecutive launches overlap (programmatic dependent launch): a call's kernels start their prologue while the
    previous call drains.  Many calls into the same backwards ring, accumulating peaks, an average that reads earlier
    calls' rows -- queued WITHOUT any host synchronisation in between, checked once at the end."""
    import torch
    import rfanalyzer_b200 as rfa
    L = 6
    L_o = oracle.lib()
    proc = L_o.orc_fftproc_new(ring, 1)
    plan = rfa.SpectrumPlan(gpu_ctx, fmt, n, avg_len=L)
    counts = [3, 1, 2, 7, 1, 1, 5, 2, 9, 4, 1, 3, 2, 2, 6, 1, 8, 3, 1, 2]
    inputs, first = [], 0
    for frames in counts:
        iq = oracle.synth_iq(fmt, n * frames, first=first)
        first += n * frames
        r, _, _ = oracle.spectrum_run(fmt, iq, n, 0)
        for k in range(frames):
            L_o.orc_fftproc_push(proc, np.ascontiguousarray(r[k]), n, 100_000_000, 20_000_000)
        inputs.append(iq)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d_in = [torch.from_numpy(iq).cuda() for iq in inputs]
        d_ring = torch.full((ring, n), -9999.0, dtype=torch.float32, device="cuda")
        d_peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
        d_avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        gpu_ctx.sync()
        write_index, history = 0, 0
        for call, frames in enumerate(counts):      # no synchronisation inside this loop
            plan.process(d_in[call], frames, rows=d_ring, peaks=d_peaks, avg=d_avg, row0=write_index, row_step=-1,
                         ring_rows=ring, history_rows=history, peaks_accumulate=call > 0)
            write_index = (write_index - frames) % ring
            history = min(ring, history + frames)
        gpu_ctx.sync()
    ring_ref = np.stack([np.ctypeslib.as_array(L_o.orc_fftproc_row(proc, i), shape=(n,)) for i in range(ring)])
    assert np.abs(d_ring.cpu().numpy() - ring_ref).max() < DB_TOL
    peaks_ref = np.ctypeslib.as_array(L_o.orc_fftproc_peaks(proc), shape=(n,))
    assert np.abs(d_peaks.cpu().numpy() - peaks_ref).max() < DB_TOL
    avg_ref = np.empty(n, np.float32)
    L_o.orc_time_average(proc, L, avg_ref)
    assert np.abs(d_avg.cpu().numpy() - avg_ref).max() < DB_TOL
    L_o.orc_fftproc_free(proc)
