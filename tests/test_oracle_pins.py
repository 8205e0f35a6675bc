"""Pins the CPU oracle (oracle/) to the reference: its own golden vectors and property tests
(app/src/androidTest/...), and the reference's native code compiled in place (oracle/_ref)."""
import ctypes as C
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _two_tone(n, sr=1000):
    # ApplicationTest.kt:33-38: cos(2*PI*f*i / sampleRate.toFloat()).toFloat(), summed in float
    i = np.arange(n)
    fs = float(np.float32(sr))
    c = lambda f: np.cos(2 * np.pi * f * i / fs).astype(np.float32)
    s = lambda f: np.sin(2 * np.pi * f * i / fs).astype(np.float32)
    return (c(50) + c(200)).astype(np.float32), (s(50) + s(200)).astype(np.float32)


def test_fir_golden_vector_1(oracle):
    """ApplicationTest.testFirFilter (:20-127): 55 taps, decimate by 4, 32 outputs, |d| <= 1e-9."""
    O, L = oracle, oracle.lib()
    k = np.load(os.path.join(GOLD, "fir_kat.npz"))
    re, im = _two_tone(128)
    f = L.orc_fir_lowpass(4, 1.0, 1000.0, 100.0, 50.0, 60.0)
    assert L.orc_fir_ntaps(f) == 55
    pin = O.PacketView(128).load(re, im, 1000)
    pout = O.PacketView(32)
    assert L.orc_fir_filter(f, pin.p, pout.p, 0, 128) == 128
    assert pout.size == 32 and pout.sampleRate == 250
    assert np.abs(pout.out_re().astype(np.float64) - k["re1"]).max() <= 1e-9
    assert np.abs(pout.out_im().astype(np.float64) - k["im1"]).max() <= 1e-9
    assert np.array_equal(pout.out_re(), k["re1"]) and np.array_equal(pout.out_im(), k["im1"])  # bit exact
    L.orc_fir_free(f)


def test_fir_golden_vector_2(oracle):
    """ApplicationTest.testFirFilter2 (:129-176): decimation 1 yields N-1 = 63 outputs."""
    O, L = oracle, oracle.lib()
    k = np.load(os.path.join(GOLD, "fir_kat.npz"))
    re, im = _two_tone(64)
    f = L.orc_fir_lowpass(1, 1.0, 1000.0, 100.0, 100.0, 40.0)
    pin = O.PacketView(64).load(re, im, 1000)
    pout = O.PacketView(64)
    assert L.orc_fir_filter(f, pin.p, pout.p, 0, 64) == 64
    assert pout.size == 63
    assert np.array_equal(pout.out_re(), k["re2"]) and np.array_equal(pout.out_im(), k["im2"])
    L.orc_fir_free(f)


def test_fir_output_capacity_stops_early(oracle):
    """FirFilter.kt:80-84: a full output packet returns the number of samples consumed so far."""
    O, L = oracle, oracle.lib()
    re, im = _two_tone(128)
    f = L.orc_fir_lowpass(4, 1.0, 1000.0, 100.0, 50.0, 60.0)
    pin = O.PacketView(128).load(re, im, 1000)
    pout = O.PacketView(10)
    consumed = L.orc_fir_filter(f, pin.p, pout.p, 0, 128)
    assert pout.size == 10 and consumed == 4 * 10 + 3  # 11th output would be emitted at input index 43
    L.orc_fir_free(f)


def test_limit_denominator(oracle):
    """RationalResamplerTest.testApproximation (:101-115) and a stride of testMaxApproximationError."""
    num, den = oracle.limit_denominator(2500101, 250000, 10000)
    assert den <= 10000 and abs(2500101 / 250000 - num / den) < 1e-4
    worst = 0.0
    for n in range(1_000_000, 10_000_001, 7919):
        a, b = oracle.limit_denominator(n, 250000, 10000)
        assert b <= 10000
        worst = max(worst, abs(n / 250000 - a / b))
    assert worst < 1e-4
    assert oracle.limit_denominator(384000, 2400000) == (4, 25)      # C2
    assert oracle.limit_denominator(96000, 10000000) == (6, 625)     # C4 NFM/SSB
    assert oracle.limit_denominator(48000, 10000000) == (3, 625)     # C4 CW


def test_resampler_geometry(oracle):
    """Taps per phase quoted in SURVEY.md 8(a) a19 / BASELINE.md."""
    L = oracle.lib()
    for (i, d, nt) in ((4, 25, 206), (6, 625, 501), (3, 625, 501), (12, 625, 501)):
        r = L.orc_resampler_new(i, d, None, 0, 0.4, 500)
        assert (L.orc_resampler_interp(r), L.orc_resampler_decim(r)) == (i, d)
        assert L.orc_resampler_taps_per_phase(r) == nt
        L.orc_resampler_free(r)


def test_resampler_round_trip(oracle):
    """RationalResamplerTest.testResamplerRoundTrip (:17-99): 11/17 then 17/11, RMSE < 0.05 at delay 51."""
    O, L = oracle, oracle.lib()
    n, sr = 2000, 48000
    t = (np.arange(n) / np.float32(sr)).astype(np.float32)
    re = np.cos(2.0 * np.pi * 100.0 * t.astype(np.float64)).astype(np.float32)
    im = np.sin(2.0 * np.pi * 100.0 * t.astype(np.float64)).astype(np.float32)
    pin = O.PacketView(n).load(re, im, sr)
    down = L.orc_resampler_new(11, 17, None, 0, 0.4, 0)
    tmp = O.PacketView(n * 11 // 17 + 100)
    assert L.orc_resampler_resample(down, pin.p, tmp.p, 0, n) == n
    up = L.orc_resampler_new(17, 11, None, 0, 0.4, 0)
    out = O.PacketView(n + 100)
    assert L.orc_resampler_resample(up, tmp.p, out.p, 0, tmp.size) == tmp.size
    m = min(n, out.size)
    delay = 51
    idx = np.arange(m - 1000, m - delay)
    dr = re[idx] - out.re[idx + delay]
    di = im[idx] - out.im[idx + delay]
    assert np.sqrt(np.mean(dr.astype(np.float64) ** 2 + di.astype(np.float64) ** 2)) < 0.05
    L.orc_resampler_free(down)
    L.orc_resampler_free(up)


def test_resampler_matches_decimator(oracle):
    """ResamplerTest.testResamplerMatchesDecimator (:21-115): 48k -> 12k in 1024-sample packets,
    legacy Decimator (Decimator.java:176-191) vs Resampler, MSE < 0.003 at delay 11."""
    O, L = oracle, oracle.lib()
    in_rate, out_rate, ps = 48000, 12000, 1024
    n = in_rate
    t = np.arange(n) / in_rate
    re = np.cos(2.0 * np.pi * 100.0 * t).astype(np.float32)
    im = np.sin(2.0 * np.pi * 100.0 * t).astype(np.float32)
    dec = L.orc_fir_lowpass(4, 1.0, float(in_rate), out_rate * 0.75, out_rate * 0.25, 60.0)
    i, d = oracle.limit_denominator(out_rate, in_rate, 10000)
    rs = L.orc_resampler_new(i, d, None, 0, 0.4, 500)
    o1r, o1i, o2r, o2i = [], [], [], []
    for pos in range(0, n, ps):
        m = min(ps, n - pos)
        pin = O.PacketView(m).load(re[pos:pos + m], im[pos:pos + m], in_rate)
        a, b = O.PacketView(ps), O.PacketView(ps)
        L.orc_fir_filter(dec, pin.p, a.p, 0, m)
        L.orc_resampler_resample(rs, pin.p, b.p, 0, m)
        o1r.append(a.out_re()); o1i.append(a.out_im()); o2r.append(b.out_re()); o2i.append(b.out_im())
    o1r, o1i, o2r, o2i = map(np.concatenate, (o1r, o1i, o2r, o2i))
    k = min(len(o1r), len(o2r))
    delay = 11
    dr = o1r[:k - delay] - o2r[delay:k]
    di = o1i[:k - delay] - o2i[delay:k]
    assert np.mean(dr.astype(np.float64) ** 2 + di.astype(np.float64) ** 2) < 0.003
    L.orc_fir_free(dec)
    L.orc_resampler_free(rs)


def test_resampler_is_a_stream(oracle):
    """Packetisation must not change the resampler's output (state carried in ctr / delay line)."""
    O, L = oracle, oracle.lib()
    rng = np.random.default_rng(3)
    n = 5000
    re = rng.standard_normal(n).astype(np.float32)
    im = rng.standard_normal(n).astype(np.float32)
    whole = L.orc_resampler_new(4, 25, None, 0, 0.4, 500)
    pin = O.PacketView(n).load(re, im, 2400000)
    pout = O.PacketView(n)
    L.orc_resampler_resample(whole, pin.p, pout.p, 0, n)
    ref_re = pout.out_re()
    assert len(ref_re) == -(-n * 4 // 25)  # ceil(n*I/D)
    parts = L.orc_resampler_new(4, 25, None, 0, 0.4, 500)
    got = []
    for pos in range(0, n, 777):
        m = min(777, n - pos)
        a = O.PacketView(m).load(re[pos:pos + m], im[pos:pos + m], 2400000)
        b = O.PacketView(m)
        assert L.orc_resampler_resample(parts, a.p, b.p, 0, m) == m
        got.append(b.out_re())
    assert np.array_equal(np.concatenate(got), ref_re)
    L.orc_resampler_free(whole)
    L.orc_resampler_free(parts)


def test_converter_luts(oracle):
    """Signed8BitIQConverter.java:48-50, Unsigned8BitIQConverter.java:48-50, Signed16BitIQConverter.kt:46-57."""
    L = oracle.lib()
    n = C.c_int()
    for fmt, expect in ((0, [(i - 128) / 128.0 for i in range(256)]),
                        (1, [np.float32(np.float32(np.float32(i) - np.float32(127.4)) / np.float32(128.0)) for i in range(256)]),
                        (2, [np.int16(np.uint16(u).astype(np.int16)) / 32768.0 for u in range(65536)])):
        c = L.orc_converter_new(fmt)
        p = L.orc_converter_lut(c, C.byref(n))
        lut = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        assert n.value == len(expect)
        assert np.array_equal(lut, np.array(expect, dtype=np.float32))
        L.orc_converter_free(c)


@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_converter_fill_and_mix(oracle, fmt):
    """fill appends at size(), stops at capacity; mix rotates by e^{+j 2 pi f t / fs} with the
    (mix == 0 or fs/|mix| > 500) -> mix += fs rule and a table of calcOptimalCosineLength entries."""
    O, L = oracle, oracle.lib()
    bps = O.BYTES_PER_SAMPLE[fmt]
    iq = O.synth_iq(fmt, 1000)
    c = L.orc_converter_new(fmt)
    L.orc_converter_set_sample_rate(c, 2400000)
    L.orc_converter_set_frequency(c, 100_000_000)
    sp = O.PacketView(600)
    assert L.orc_converter_fill(c, iq[: 400 * bps], 400 * bps, sp.p) == 400
    assert L.orc_converter_fill(c, iq[400 * bps:], 600 * bps, sp.p) == 200  # capacity reached
    assert sp.size == 600 and sp.sampleRate == 2400000 and sp.frequency == 100_000_000
    plain_re, plain_im = sp.out_re(), sp.out_im()
    mp = O.PacketView(600)
    assert L.orc_converter_mix(c, iq, 1000 * bps, mp.p, 100_250_000) == 600
    assert mp.frequency == 100_250_000
    f = L.orc_converter_nco_freq(c)
    ln = L.orc_converter_nco_len(c)
    assert f == -250000 and ln == oracle.lib().orc_calc_optimal_cosine_length(2400000, -250000) and 0 < ln <= 500
    t = np.arange(600) % ln
    rot = np.exp(2j * np.pi * f * t / 2400000.0)
    want = (plain_re + 1j * plain_im) * rot
    got = mp.out_re() + 1j * mp.out_im()
    assert np.abs(got - want).max() < 2e-6
    assert L.orc_converter_nco_index(c) == 600 % ln
    # tiny mix frequency: table would exceed 500 entries, so the sample rate is added
    tiny = O.PacketView(4)
    L.orc_converter_mix(c, iq, 4 * bps, tiny.p, 100_000_100)
    assert L.orc_converter_nco_freq(c) == -100 + 2400000
    L.orc_converter_free(c)


@pytest.mark.parametrize("n", [64, 1024, 4096, 65536])
def test_fft_restatement_vs_double(oracle, n):
    L = oracle.lib()
    x = np.random.default_rng(n).standard_normal(2 * n).astype(np.float32)
    a = np.empty(2 * n, np.float32)
    L.orc_fft_c2c_f32(x, a, n)
    ref = np.fft.fft(x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64))
    got = a[0::2] + 1j * a[1::2]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-6


@pytest.mark.parametrize("n", [1024, 4096, 16384, 65536])
def test_restatement_vs_compiled_reference(oracle, n):
    """The reference's own pffft.c + nativedsp.cpp (oracle/_ref) pin the restated FFT/log-mag."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L, R = oracle.lib(), oracle.ref()
    assert R.ref_pffft_simd_size() == 4
    x = np.random.default_rng(n + 1).standard_normal(2 * n).astype(np.float32)
    a, b = np.empty(2 * n, np.float32), np.empty(2 * n, np.float32)
    L.orc_fft_c2c_f32(x, a, n)
    R.ref_perform_fft(x, b, 2 * n)
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-6
    ma, mb = np.empty(n, np.float32), np.empty(n, np.float32)
    L.orc_fft_logmag(x, ma, n)
    R.ref_perform_fft_logmag(x, mb, 2 * n)
    assert np.abs(ma - mb).max() < 0.01


def test_spectrum_chain_vs_reference_golden(oracle):
    """orc_spectrum_run against rows produced by the compiled reference (tests/golden)."""
    g = np.load(os.path.join(GOLD, "spectrum_ref.npz"))
    for fmt, name in ((0, "s8"), (1, "u8"), (2, "s16")):
        for n in (1024, 4096):
            iq = oracle.synth_iq(fmt, n * 2, first=12345)
            rows, peaks, avg = oracle.spectrum_run(fmt, iq, n, 1)
            assert np.abs(rows - g[f"{name}_{n}_rows"]).max() < 0.01
            assert np.abs(peaks - g[f"{name}_{n}_peaks"]).max() < 0.01
            assert np.abs(avg - g[f"{name}_{n}_avg"]).max() < 0.01


def test_nativedsp_window(oracle):
    w = oracle.nativedsp_window(4096)
    i = np.arange(4096)
    ref = (0.42 - 0.5 * np.cos(2 * np.pi * i / 4095) + 0.08 * np.cos(4 * np.pi * i / 4095)).astype(np.float32)
    assert np.array_equal(w, ref)


def test_fftproc_ring_peaks_average(oracle):
    """FftProcessor.kt:178-245 ring runs backwards, peaks reset on retune; AnalyzerSurface.kt:710-714."""
    L = oracle.lib()
    n = 64
    p = L.orc_fftproc_new(300, 1)
    rng = np.random.default_rng(0)
    rows = (rng.standard_normal((5, n)) * 10 - 50).astype(np.float32)
    idxs = [L.orc_fftproc_push(p, np.ascontiguousarray(rows[k]), n, 100_000_000, 1_000_000) for k in range(5)]
    assert idxs == [0, 299, 298, 297, 296]
    assert L.orc_fftproc_read_index(p) == 296 and L.orc_fftproc_write_index(p) == 295
    peaks = np.ctypeslib.as_array(L.orc_fftproc_peaks(p), shape=(n,)).copy()
    assert np.array_equal(peaks, rows.max(axis=0))
    avg = np.empty(n, np.float32)
    L.orc_time_average(p, 2, avg)
    s = np.zeros(n, np.float32)
    for k in (4, 3, 2):
        s = (s + rows[k]).astype(np.float32)
    assert np.array_equal(avg, (s / np.float32(3)).astype(np.float32))
    # retune by +100 kHz: history shifts by -6 bins (100e3 * 64/1e6 = 6.4 -> toInt), peaks reset
    L.orc_fftproc_push(p, np.ascontiguousarray(rows[0]), n, 100_100_000, 1_000_000)
    old = np.ctypeslib.as_array(L.orc_fftproc_row(p, 296), shape=(n,)).copy()
    assert np.array_equal(old[: n - 6], rows[4][6:]) and np.all(old[n - 6:] == -9999.0)
    peaks = np.ctypeslib.as_array(L.orc_fftproc_peaks(p), shape=(n,)).copy()
    assert np.array_equal(peaks, rows[0])
    L.orc_fftproc_free(p)


def test_synth_is_deterministic_and_seekable(oracle):
    a = oracle.synth_iq(0, 5000)
    b = np.concatenate([oracle.synth_iq(0, 1234), oracle.synth_iq(0, 5000 - 1234, first=1234)])
    assert np.array_equal(a, b)
    s16 = oracle.synth_iq(2, 1000).view(np.int16)
    assert s16.max() < 32767 and s16.min() > -32768 and np.abs(s16).max() > 10000


@pytest.mark.parametrize("mode,fs,cw,fmt,packet", [(3, 2400000, 100000, 1, 8192), (2, 10000000, 10000, 2, 65536),
                                                   (5, 10000000, 2800, 2, 65536), (6, 10000000, 300, 2, 65536),
                                                   (1, 10000000, 8000, 2, 65536)])
def test_chain_produces_audio(oracle, mode, fs, cw, fmt, packet):
    """orc_chain_run output rate: 48 kHz audio for every mode (Demodulator.kt:53-62, AudioSink.java:215-237)."""
    n = packet * 6
    iq = oracle.synth_iq(fmt, n)
    audio = oracle.chain_run(fmt, iq, fs, 100_000_000, 100_000_000 + fs // 10, mode, cw, packet)
    expect = n * 48000.0 / fs
    assert abs(len(audio) - expect) <= 8 + 0.02 * expect
    assert np.all(np.isfinite(audio))


def test_iqconverter_restatement_equals_the_compiled_reference(oracle):
    """iqconverter_int16.c (libairspy) compiled in place vs the C restatement: identical int16 output for noise,
    full-scale square waves (int16 wrap-around), silence and constant input, across ragged calls."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref is not built (no reference tree on this box)")
    rng = np.random.default_rng(5)
    k = oracle.airspy_hb_kernel()
    assert len(k) == 47 and k[23] == 16384 and k[1] == 0
    n = 60_000
    t = np.arange(n)
    sigs = [rng.integers(-32768, 32768, n), (rng.integers(0, 4096, n) - 2048) << 4, np.where(t % 5 < 2, -32768, 32767),
            np.zeros(n), np.full(n, -777)]
    for sig in sigs:
        x = sig.astype(np.int16)
        a, b = oracle.IqConverterInt16(k, use_ref=False), oracle.IqConverterInt16(k, use_ref=True)
        for lo, hi in ((0, 4), (4, 20000), (20000, 20008), (20008, n)):
            assert np.array_equal(a.process(x[lo:hi].copy()), b.process(x[lo:hi].copy()))
