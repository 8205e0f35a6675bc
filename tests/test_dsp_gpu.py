"""FirFilter / ComplexFirFilter / RationalResampler / Demodulator / AudioSink and the whole IQ->audio
chain on the GPU (through the C ABI and its host-side mirror) against the oracle.

RFA_SUM_EXACT reproduces the JVM's float32 arithmetic: results are compared BIT FOR BIT.
RFA_SUM_FMA (the fast default) is held to BASELINE.json's tolerance for demodulated audio:
1e-4 of the signal's peak (0.01 dB is 1.15e-3 relative)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def two_tone(n, sr=1000):
    i = np.arange(n)
    fs = float(np.float32(sr))
    c = lambda f: np.cos(2 * np.pi * f * i / fs).astype(np.float32)
    s = lambda f: np.sin(2 * np.pi * f * i / fs).astype(np.float32)
    return (c(50) + c(200)).astype(np.float32), (s(50) + s(200)).astype(np.float32)


def test_fir_filter_reference_known_answers(gpu_ctx):
    """ApplicationTest.testFirFilter / testFirFilter2 (ApplicationTest.kt:20-176), run as written."""
    import rfanalyzer_b200 as rfa
    k = np.load(os.path.join(GOLD, "fir_kat.npz"))
    re, im = two_tone(128)
    inp = rfa.SamplePacket(re, im, 0, 1000)
    out = rfa.SamplePacket(np.zeros(32, np.float32), np.zeros(32, np.float32), 0, 250)
    out.setSize(0)
    f = rfa.FirFilter.createLowPass(gpu_ctx, 4, 1.0, 1000.0, 100.0, 50.0, 60.0)
    assert f.numberOfTaps == 55
    assert f.filter(inp, out, 0, inp.size()) == 128
    assert out.size() == 32
    assert np.abs(out.re().astype(np.float64) - k["re1"]).max() <= 1e-9   # the reference's own assertion
    assert np.abs(out.im().astype(np.float64) - k["im1"]).max() <= 1e-9
    assert np.array_equal(out.re(), k["re1"]) and np.array_equal(out.im(), k["im1"])
    re, im = two_tone(64)
    inp = rfa.SamplePacket(re, im, 0, 1000)
    out = rfa.SamplePacket(np.zeros(64, np.float32), np.zeros(64, np.float32), 0, 250)
    out.setSize(0)
    f = rfa.FirFilter.createLowPass(gpu_ctx, 1, 1.0, 1000.0, 100.0, 100.0, 40.0)
    f.filter(inp, out, 0, inp.size())
    assert out.size() == 63
    assert np.array_equal(out.re()[:63], k["re2"]) and np.array_equal(out.im()[:63], k["im2"])


@pytest.mark.parametrize("dec", [1, 2, 4, 7])
def test_fir_streaming_state_bit_exact(gpu_ctx, oracle, dec):
    """Delay line and decimationCounter carry across packets of ragged length; a full output packet
    stops the call early with the reference's return value (FirFilter.kt:80-84)."""
    import rfanalyzer_b200 as rfa
    L = oracle.lib()
    rng = np.random.default_rng(dec)
    taps = oracle.lowpass_taps(1.0, 48000.0, 9000.0 / dec, 3000.0, 60.0)
    of = L.orc_fir_new(taps, len(taps), dec)
    gf = rfa.FirFilter(gpu_ctx, taps, dec)
    for n, cap in ((1000, 2000), (1, 10), (37, 10), (5000, 100), (3, 3), (4096, 4096)):
        re = rng.standard_normal(n).astype(np.float32)
        im = rng.standard_normal(n).astype(np.float32)
        pin = oracle.PacketView(n).load(re, im, 48000)
        pout = oracle.PacketView(cap)
        want_consumed = L.orc_fir_filter(of, pin.p, pout.p, 0, n)
        out = rfa.SamplePacket(cap)
        got_consumed = gf.filter(rfa.SamplePacket(re, im, 0, 48000), out, 0, n)
        assert got_consumed == want_consumed and out.size() == pout.size
        assert np.array_equal(out.re()[: out.size()], pout.out_re())
        assert np.array_equal(out.im()[: out.size()], pout.out_im())
    L.orc_fir_free(of)


def test_filter_real_and_fma_mode(gpu_ctx, oracle):
    import rfanalyzer_b200 as rfa
    L = oracle.lib()
    x = np.random.default_rng(1).standard_normal(20000).astype(np.float32)
    of = L.orc_fir_lowpass(2, 1.0, 1.0, 0.1, 0.15, 30.0)
    pin = oracle.PacketView(20000).load(x, None, 96000)
    pout = oracle.PacketView(20000)
    L.orc_fir_filter_real(of, pin.p, pout.p, 0, 20000)
    for flags, exact in ((rfa.SUM_EXACT, True), (rfa.SUM_FMA, False)):
        gf = rfa.FirFilter.createLowPass(gpu_ctx, 2, 1.0, 1.0, 0.1, 0.15, 30.0, flags)
        assert gf.numberOfTaps == 9
        out = rfa.SamplePacket(20000)
        gf.filterReal(rfa.SamplePacket(x, np.zeros_like(x), 0, 96000), out, 0, 20000)
        assert out.size() == pout.size == 10000 and out.sampleRate == 48000
        if exact:
            assert np.array_equal(out.re()[:10000], pout.out_re())
        else:
            assert np.abs(out.re()[:10000] - pout.out_re()).max() < 1e-5 * np.abs(pout.out_re()).max()
    L.orc_fir_free(of)


@pytest.mark.parametrize("dec,lo,hi,fs", [(2, 200.0, 2800.0, 96000.0), (2, -2800.0, -200.0, 96000.0), (1, 600.0, 900.0, 48000.0)])
def test_complex_fir_bit_exact(gpu_ctx, oracle, dec, lo, hi, fs):
    """ComplexFirFilter.createBandPass + filter (the SSB / CW band-pass), two consecutive packets."""
    import rfanalyzer_b200 as rfa
    L = oracle.lib()
    of = L.orc_cfir_bandpass(dec, 1.0, fs, lo, hi, fs * 0.01, 40.0)
    gf = rfa.ComplexFirFilter.createBandPass(gpu_ctx, dec, 1.0, fs, lo, hi, float(np.float32(fs) * np.float32(0.01)), 40.0)
    assert gf.getNumberOfTaps() == L.orc_cfir_ntaps(of) == 181
    rng = np.random.default_rng(7)
    for n in (3000, 1111):
        re = rng.standard_normal(n).astype(np.float32)
        im = rng.standard_normal(n).astype(np.float32)
        pin = oracle.PacketView(n).load(re, im, int(fs))
        pout = oracle.PacketView(n)
        L.orc_cfir_filter(of, pin.p, pout.p, 0, n)
        out = rfa.SamplePacket(n)
        gf.filter(rfa.SamplePacket(re, im, 0, int(fs)), out, 0, n)
        assert out.size() == pout.size
        assert np.array_equal(out.re()[: out.size()], pout.out_re()) and np.array_equal(out.im()[: out.size()], pout.out_im())
    L.orc_cfir_free(of)


def test_resampler_round_trip(gpu_ctx):
    """RationalResamplerTest.testResamplerRoundTrip (RationalResamplerTest.kt:17-99) as written."""
    import rfanalyzer_b200 as rfa
    interpolation, decimation, sampleRate, numSamples = 11, 17, 48000, 2000
    t = (np.arange(numSamples) / np.float32(sampleRate)).astype(np.float32).astype(np.float64)
    reIn = np.cos(2.0 * np.pi * 100.0 * t).astype(np.float32)
    imIn = np.sin(2.0 * np.pi * 100.0 * t).astype(np.float32)
    inPacket = rfa.SamplePacket(reIn, imIn, 0, sampleRate)
    down = rfa.RationalResampler(gpu_ctx, interpolation, decimation)
    tmp = rfa.SamplePacket(numSamples * interpolation // decimation + 100)
    assert down.resample(inPacket, tmp, 0, numSamples) == numSamples
    up = rfa.RationalResampler(gpu_ctx, decimation, interpolation)
    out = rfa.SamplePacket(numSamples + 100)
    assert up.resample(tmp, out, 0, tmp.size()) == tmp.size()
    minLen, delay = min(numSamples, out.size()), 51
    idx = np.arange(minLen - 1000, minLen - delay)
    dr = reIn[idx] - out.re()[idx + delay]
    di = imIn[idx] - out.im()[idx + delay]
    assert np.sqrt(np.mean(dr.astype(np.float64) ** 2 + di.astype(np.float64) ** 2)) < 0.05
    assert rfa.RationalResampler.limitDenominator(2500101, 250000, 10000)[1] <= 10000


@pytest.mark.parametrize("i,d,maxtaps", [(4, 25, 500), (6, 625, 500), (11, 17, 0), (17, 11, 0), (3, 2, 0)])
def test_resampler_streaming_bit_exact(gpu_ctx, oracle, i, d, maxtaps):
    """ctr / delay-line state across ragged packets, and the output-capacity stop (RationalResampler.kt:119)."""
    import rfanalyzer_b200 as rfa
    L = oracle.lib()
    orr = L.orc_resampler_new(i, d, None, 0, 0.4, maxtaps)
    grr = rfa.RationalResampler(gpu_ctx, i, d, maxTaps=maxtaps)
    assert grr.tapsPerPhase == L.orc_resampler_taps_per_phase(orr)
    rng = np.random.default_rng(i * 100 + d)
    for n, cap in ((8192, 8192 * 2), (1, 4), (999, 4000), (20000, 50), (7, 100), (30011, 60000)):
        re = rng.standard_normal(n).astype(np.float32)
        im = rng.standard_normal(n).astype(np.float32)
        pin = oracle.PacketView(n).load(re, im, 2400000)
        pout = oracle.PacketView(cap)
        want = L.orc_resampler_resample(orr, pin.p, pout.p, 0, n)
        out = rfa.SamplePacket(cap)
        got = grr.resample(rfa.SamplePacket(re, im, 0, 2400000), out, 0, n)
        assert got == want and out.size() == pout.size
        assert np.array_equal(out.re()[: out.size()], pout.out_re())
        assert np.array_equal(out.im()[: out.size()], pout.out_im())
        assert out.sampleRate == pout.sampleRate
    L.orc_resampler_free(orr)


@pytest.mark.parametrize("i,d,maxtaps", [(4, 25, 500), (1, 25, 500), (2, 5, 500), (6, 625, 500), (12, 625, 500), (3, 625, 500), (6, 125, 500), (12, 3125, 500),
                                         (11, 17, 0), (17, 11, 0), (3, 2, 0), (4, 50, 500),
                                         # stripe kernel corner cases: one phase (every task split six ways), more tasks than warps, odd I
                                         (1, 400, 500), (16, 625, 500), (5, 312, 500), (7, 2000, 500)])
def test_resampler_fast_paths_vs_oracle(gpu_ctx, oracle, i, d, maxtaps):
    """RFA_SUM_FMA selects the register-tiled kernel (taps per phase > D) or the lanes-per-output kernel:
    same counters and state as the exact kernel, samples within the demodulated-audio tolerance
    (the additions are reordered, every product is the same)."""
    import rfanalyzer_b200 as rfa
    L = oracle.lib()
    orr = L.orc_resampler_new(i, d, None, 0, 0.4, maxtaps)
    grr = rfa.RationalResampler(gpu_ctx, i, d, maxTaps=maxtaps, flags=rfa.SUM_FMA)
    rng = np.random.default_rng(i * 1000 + d)
    for n, cap in ((8192, 8192 * 2), (1, 4), (999, 4000), (20000, 50), (7, 100), (70011, 140000), (65536, 140000)):
        re = rng.standard_normal(n).astype(np.float32)
        im = rng.standard_normal(n).astype(np.float32)
        pin = oracle.PacketView(n).load(re, im, 2400000)
        pout = oracle.PacketView(cap)
        want = L.orc_resampler_resample(orr, pin.p, pout.p, 0, n)
        out = rfa.SamplePacket(cap)
        got = grr.resample(rfa.SamplePacket(re, im, 0, 2400000), out, 0, n)
        assert got == want and out.size() == pout.size
        if out.size():
            wr, wi = pout.out_re(), pout.out_im()
            peak = max(np.abs(wr).max(), np.abs(wi).max(), 1e-30)
            assert np.abs(out.re()[: out.size()] - wr).max() <= 1e-5 * peak + 1e-7
            assert np.abs(out.im()[: out.size()] - wi).max() <= 1e-5 * peak + 1e-7
    L.orc_resampler_free(orr)


def _quad_packets(oracle, mode, npackets, n):
    """complex packets at the mode's quadrature rate: a modulated carrier plus noise"""
    rng = np.random.default_rng(mode)
    rate = oracle.lib().orc_mode_quadrature_rate(mode)
    out = []
    ph = 0.0
    for p in range(npackets):
        t = (np.arange(n) + p * n) / rate
        if mode in (2, 3):
            phase = 2 * np.pi * 0.3 * np.cumsum(np.sin(2 * np.pi * 1000 * t)) + ph
            x = 0.5 * np.exp(1j * phase)
            ph = phase[-1]
        else:
            x = (0.3 + 0.2 * np.sin(2 * np.pi * 700 * t)) * np.exp(2j * np.pi * 900 * t) * (1 + 0.3 * p)
        x = x + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        out.append((x.real.astype(np.float32), x.imag.astype(np.float32)))
    return rate, out


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5, 6])
def test_demodulator_packets_vs_oracle(gpu_ctx, oracle, mode):
    """Demodulator.run body per packet (user filter, demodulate*, volume) and AudioSink.applyAudioFilter."""
    import rfanalyzer_b200 as rfa
    L = oracle.lib()
    n = 6000
    rate, packets = _quad_packets(oracle, mode, 4, n)
    od = L.orc_demod_new(n)
    L.orc_demod_set_mode(od, mode)
    L.orc_demod_set_volume(od, 0.8)
    osink = L.orc_audiosink_new(n, 48000)
    gd = rfa.Demodulator(gpu_ctx, n)
    gd.demodulationMode = mode
    gd.audioVolumeLevel = 0.8
    gsink = rfa.AudioSink(gpu_ctx, n)
    assert gd.channelWidth == L.orc_demod_channel_width(od)
    for re, im in packets:
        pin = oracle.PacketView(n).load(re, im, rate)
        pa, pf = oracle.PacketView(n), oracle.PacketView(n)
        L.orc_demod_process(od, pin.p, pa.p)
        ga, gf = rfa.SamplePacket(n), rfa.SamplePacket(n)
        gd.process(rfa.SamplePacket(re, im, 0, rate), ga)
        assert ga.size() == pa.size and ga.sampleRate == pa.sampleRate
        want, got = pa.out_re(), ga.re()[: ga.size()]
        if mode in (2, 3):  # double-precision atan2 on both sides: identical up to libm's last bit
            assert np.abs(got - want).max() <= 2e-7 * max(1.0, np.abs(want).max())
            assert np.mean(got == want) > 0.99
        else:
            assert np.array_equal(got, want)
        if pa.sampleRate > 48000:
            assert L.orc_audiosink_filter(osink, pa.p, pf.p) == 1
            ga_exact = rfa.SamplePacket(want.copy(), np.zeros_like(want), 0, pa.sampleRate)
            assert gsink.applyAudioFilter(ga_exact, gf)
            assert gf.size() == pf.size and np.array_equal(gf.re()[: gf.size()], pf.out_re())
    L.orc_demod_free(od)
    L.orc_audiosink_free(osink)


CHAINS = [  # fmt, fs, mode, width, packet samples, npackets    (BASELINE configs 2 and 4)
    (1, 2_400_000, 3, 100_000, 8192, 24),      # C2: RTL-SDR uint8 -> wFM
    (2, 10_000_000, 2, 10_000, 65536, 5),      # C4: Airspy int16 -> nFM
    (2, 10_000_000, 5, 2_800, 65536, 5),       # USB
    (2, 10_000_000, 4, 2_800, 65536, 5),       # LSB
    (2, 10_000_000, 6, 300, 65536, 5),         # CW
    (2, 10_000_000, 1, 8_000, 65536, 5),       # AM
    (0, 20_000_000, 3, 100_000, 131072, 3),    # HackRF int8 @20 Msps -> wFM (12/625, 501 taps/phase)
]


def _chain_input(oracle, rfa, fmt, fs, mode, nsamples):
    off = fs // 10
    mul = 256 if fmt == 2 else 1
    comps = [(rfa.synth_step(off / fs), 60 * mul, rfa.synth_step(1000 / fs), 3_130_000 if mode in (2, 3) else 0),
             (rfa.synth_step((off + 1200) / fs), 20 * mul, 0, 0)]
    return oracle.synth_iq(fmt, nsamples, comps=comps, noise_shift=3), 100_000_000, 100_000_000 + off


@pytest.mark.parametrize("fmt,fs,mode,width,packet,npackets", CHAINS)
@pytest.mark.parametrize("exact", [True, False])
def test_chain_vs_oracle(gpu_ctx, oracle, fmt, fs, mode, width, packet, npackets, exact):
    import rfanalyzer_b200 as rfa
    n = packet * npackets - packet // 3  # last packet is short
    iq, src, chan = _chain_input(oracle, rfa, fmt, fs, mode, n)
    want = oracle.chain_run(fmt, iq, fs, src, chan, mode, width, packet, volume=0.9)
    plan = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 0.9,
                         rfa.SUM_EXACT if exact else rfa.SUM_FMA)
    audio = np.zeros(plan.max_audio(n), np.float32)
    got_n = plan.process(iq, n, audio)
    assert got_n == len(want)
    got = audio[:got_n]
    peak = np.abs(want).max()
    assert peak > 0
    if exact and mode not in (2, 3):
        assert np.array_equal(got, want)
    elif exact:
        assert np.abs(got - want).max() <= 1e-6 * peak
    else:
        assert np.abs(got - want).max() <= 1e-4 * peak


def test_chain_streams_across_calls(gpu_ctx, oracle):
    """State (NCO index, resampler phase, delay lines, FM carry, AGC) carries across calls made on
    whole packets: two calls equal one."""
    import rfanalyzer_b200 as rfa
    fmt, fs, mode, width, packet = 1, 2_400_000, 3, 100_000, 8192
    n = packet * 12
    iq, src, chan = _chain_input(oracle, rfa, fmt, fs, mode, n)
    one = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 1.0, rfa.SUM_EXACT)
    a = np.zeros(one.max_audio(n), np.float32)
    na = one.process(iq, n, a)
    two = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 1.0, rfa.SUM_EXACT)
    b1 = np.zeros(two.max_audio(packet * 5), np.float32)
    n1 = two.process(iq[: packet * 5 * 2], packet * 5, b1)
    b2 = np.zeros(two.max_audio(packet * 7), np.float32)
    n2 = two.process(iq[packet * 5 * 2:], packet * 7, b2)
    assert n1 + n2 == na
    assert np.array_equal(np.concatenate([b1[:n1], b2[:n2]]), a[:na])


@pytest.mark.parametrize("fmt,fs,mode,width,packet,calls", [
    (2, 10_000_000, 5, 2_800, 65536, (3, 1, 4)),       # USB: band-pass /2, three calls
    (2, 10_000_000, 6, 300, 65536, (2, 5, 1)),         # CW: band-pass /1
    (2, 10_000_000, 1, 8_000, 65536, (4, 1, 3)),       # AM: power, audio decimator
    (1, 2_400_000, 1, 8_000, 64, (2, 30, 1, 50)),      # AM, calls shorter than the audio decimator's delay line
    (1, 2_400_000, 4, 2_800, 64, (2, 30, 1, 50)),      # LSB, a handful of samples per packet
    (1, 2_400_000, 5, 2_800, 512, (1100, 3)),          # USB, more packets than the in-kernel AGC scan takes
])
def test_fused_agc_tail_streams_across_calls(gpu_ctx, oracle, fmt, fs, mode, width, packet, calls):
    """RFA_SUM_FMA runs AM / SSB / CW behind the resampler in the fused kernels of chain_agc.cu: delay lines of the
    user filter, band-pass and audio decimator, the AGC maximum and the two alternating packet tables carry from call
    to call -- several calls of uneven length equal the oracle's single run (BASELINE tolerance for audio)."""
    import rfanalyzer_b200 as rfa
    n = packet * sum(calls)
    iq, src, chan = _chain_input(oracle, rfa, fmt, fs, mode, n)
    want = oracle.chain_run(fmt, iq, fs, src, chan, mode, width, packet, volume=0.8)
    plan = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 0.8, rfa.SUM_FMA)
    bps = 2 if fmt < 2 else 4
    got, pos = [], 0
    for k in calls:
        m = packet * k
        audio = np.zeros(plan.max_audio(m), np.float32)
        cnt = plan.process(iq[pos * bps:(pos + m) * bps], m, audio)
        got.append(audio[:cnt])
        pos += m
    got = np.concatenate(got)
    assert len(got) == len(want)
    ok = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), ok) and ok.sum() > len(want) // 2
    assert np.abs(got[ok] - want[ok]).max() <= 1e-4 * np.abs(want[ok]).max()


@pytest.mark.parametrize("fmt,fs,mode,width,packet,calls,exact", [
    (1, 2_400_000, 3, 100_000, 8192, (5, 1, 6), False),       # wFM: tiled resampler, both decimators
    (2, 10_000_000, 2, 10_000, 65536, (2, 1, 3), False),      # nFM: stripe resampler, first decimator
    (1, 2_400_000, 3, 100_000, 16, (3, 40, 1, 2, 60), False), # wFM, calls of two or three quadrature samples
    (1, 2_400_000, 2, 10_000, 64, (2, 30, 1, 50), True),      # nFM exact, calls shorter than the decimator's delay line
])
def test_fm_chain_streams_across_calls_with_in_kernel_delay_lines(gpu_ctx, oracle, fmt, fs, mode, width, packet, calls, exact):
    """The FM chain's delay lines (resampler, user filter, both audio decimators) slide inside the resampler and FM
    tail kernels: several calls of uneven length -- some shorter than a decimator's delay line -- equal the oracle's
    single run."""
    import rfanalyzer_b200 as rfa
    n = packet * sum(calls)
    iq, src, chan = _chain_input(oracle, rfa, fmt, fs, mode, n)
    want = oracle.chain_run(fmt, iq, fs, src, chan, mode, width, packet, volume=0.8)
    plan = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 0.8, rfa.SUM_EXACT if exact else rfa.SUM_FMA)
    bps = 2 if fmt < 2 else 4
    got, pos = [], 0
    for k in calls:
        m = packet * k
        audio = np.zeros(plan.max_audio(m), np.float32)
        cnt = plan.process(iq[pos * bps:(pos + m) * bps], m, audio)
        got.append(audio[:cnt])
        pos += m
    got = np.concatenate(got)
    assert len(got) == len(want)
    peak = np.abs(want).max()
    assert np.abs(got - want).max() <= (1e-6 if exact else 1e-4) * peak


def test_chain_device_buffers(gpu_ctx, oracle):
    import torch
    import rfanalyzer_b200 as rfa
    fmt, fs, mode, width, packet = 2, 10_000_000, 2, 10_000, 65536
    n = packet * 4
    iq, src, chan = _chain_input(oracle, rfa, fmt, fs, mode, n)
    want = oracle.chain_run(fmt, iq, fs, src, chan, mode, width, packet)
    plan = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet)
    assert (plan.interpolation, plan.decimation, plan.taps_per_phase) == (6, 625, 501)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        audio = torch.zeros(plan.max_audio(n), dtype=torch.float32, device="cuda")
        got_n = plan.process(d, n, audio)
        gpu_ctx.sync()
    assert got_n == len(want)
    assert np.abs(audio[:got_n].cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max()


def test_chain_rejects_upsampling_and_bad_modes(gpu_ctx):
    import rfanalyzer_b200 as rfa
    with pytest.raises(rfa.RfaError):
        rfa.ChainPlan(gpu_ctx, 1, 250_000, 0, 0, 3, 100_000, 8192)   # below the wFM quadrature rate
    with pytest.raises(rfa.RfaError):
        rfa.ChainPlan(gpu_ctx, 1, 2_400_000, 0, 0, 0, 0, 8192)       # OFF is not a demodulator


@pytest.mark.parametrize("fmt,fs,mode,width,packet,npackets,world,exact", [
    (1, 2_400_000, 3, 100_000, 8192, 23, 3, True),     # RTL-SDR wFM: bit-exact across shards
    (1, 2_400_000, 3, 100_000, 8192, 23, 4, False),    # the fast kernels are deterministic per output too
    (2, 10_000_000, 2, 10_000, 65536, 7, 2, True),     # Airspy nFM
    (1, 2_400_000, 1, 8_000, 1024, 1500, 3, True),     # AM: AGC memory decays over the 256-packet halo
    (2, 10_000_000, 5, 2_800, 4096, 1200, 2, False),   # USB
])
def test_time_sharded_chain_matches_the_sequential_run(gpu_ctx, oracle, fmt, fs, mode, width, packet, npackets, world, exact):
    """SURVEY.md 8(e), demodulation path: packet-aligned segments, seek + warm-up halo, no collective.
    Every 'rank' is run here one after the other on the same GPU."""
    import rfanalyzer_b200 as rfa
    from rfanalyzer_b200.sharding import ShardedChain
    n = packet * npackets - packet // 3
    iq, src, chan = _chain_input(oracle, rfa, fmt, fs, mode, n)
    flags = rfa.SUM_EXACT if exact else rfa.SUM_FMA
    bps = rfa.BYTES_PER_SAMPLE[fmt]
    whole = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 0.9, flags)
    want = np.zeros(whole.max_audio(n), np.float32)
    want = want[:whole.process(iq, n, want)]
    pieces, expect_index = [], 0
    for rank in range(world):
        plan = rfa.ChainPlan(gpu_ctx, fmt, fs, src, chan, mode, width, packet, 0.9, flags)
        sc = ShardedChain(plan, rank, world)
        halo_start, first, cnt = sc.segment(n)
        audio = np.zeros(plan.max_audio(max(first - halo_start, cnt)), np.float32)
        index, got = sc.process(iq[halo_start * bps:(first + cnt) * bps], n, audio)
        assert index == expect_index      # the closed-form counters agree with the streamed ones
        expect_index += got
        pieces.append(audio[:got].copy())
    got_all = np.concatenate(pieces)
    assert len(got_all) == len(want)
    if mode in (2, 3):
        assert np.array_equal(got_all, want)
    else:
        # the very first packets of a recording divide by an AGC maximum of zero (Demodulator.kt:299-302 does
        # too): those samples are +-inf / NaN in both runs and must sit in the same places
        fin = np.isfinite(want)
        assert np.array_equal(fin, np.isfinite(got_all)) and fin.sum() > 0.9 * len(want)
        assert np.abs(got_all[fin] - want[fin]).max() <= 1e-4 * np.abs(want[fin]).max()


@pytest.mark.parametrize("fmt,fs,mode,width,packet", [(1, 2_400_000, 3, 100_000, 8192), (0, 20_000_000, 3, 100_000, 16384),
                                                      (2, 10_000_000, 2, 10_000, 65536)])
def test_fast_staging_is_bit_identical_to_the_general_staging(gpu_ctx, oracle, fmt, fs, mode, width, packet):
    """The tiled and stripe resamplers stage 8-byte aligned integer IQ with a pair-per-thread path (magic-number
    conversion, pre-scaled NCO table, fir.cu stage_span_pairs) and everything else -- history, unaligned buffers --
    with the general one.  Both must produce the same BITS for a sample (else the audio of a stream would depend on
    how it is cut into calls and tiles): the same stream from an aligned and from a 2- / 4-byte shifted device buffer.
    (This is the test that caught ptxas contracting mul.rn.f32x2 + add.rn.f32x2 into FFMA2.)"""
    import torch
    import rfanalyzer_b200 as rfa
    n = packet * 9
    bps = rfa.BYTES_PER_SAMPLE[fmt]
    iq = oracle.synth_iq(fmt, n).view(np.uint8)
    outs = []
    with torch.cuda.stream(gpu_ctx.torch_stream):
        for off in (0, bps):
            buf = torch.zeros(len(iq) + 16, dtype=torch.uint8, device="cuda")
            buf[off:off + len(iq)] = torch.from_numpy(iq).cuda()
            plan = rfa.ChainPlan(gpu_ctx, fmt, fs, 100_000_000, 100_000_000 + fs // 10, mode, width, packet, 1.0, rfa.SUM_FMA)
            audio = torch.zeros(plan.max_audio(n), dtype=torch.float32, device="cuda")
            got = plan.process(buf[off:off + len(iq)], n, audio)
            gpu_ctx.sync()
            outs.append(audio[:got].cpu().numpy())
    assert len(outs[0]) > 100 and np.array_equal(outs[0], outs[1])
