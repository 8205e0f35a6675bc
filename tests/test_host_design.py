"""Host-side design functions of the product (librfa_b200's tables and taps) against the oracle:
they restate the same reference formulas independently and must agree bit for bit."""
import ctypes as C

import numpy as np
import pytest

from rfanalyzer_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_fft_window_bit_exact(lib, oracle):
    for n in (16, 1024, 4096, 65536):
        w = np.empty(n, np.float32)
        assert lib.rfa_make_window(_lib.WIN_BLACKMAN_REF, n, w.ctypes.data) == 0
        assert np.array_equal(w, oracle.nativedsp_window(n))
    h = np.empty(8, np.float32)
    lib.rfa_make_window(_lib.WIN_HANN, 8, h.ctypes.data)
    assert np.allclose(h, 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(8) / 7), atol=1e-7)


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("fs,mix", [(2400000, -250000), (2400000, 100), (10000000, 1234567), (20000000, 0),
                                    (2400000, -2399999), (1000000, 333333)])
def test_nco_design_bit_exact(lib, oracle, fmt, fs, mix):
    """generateMixerLookupTable + calcOptimalCosineLength, incl. the 8-bit / 16-bit angle formulas."""
    L = oracle.lib()
    c = L.orc_converter_new(fmt)
    L.orc_converter_set_sample_rate(c, fs)
    L.orc_converter_set_frequency(c, 100_000_000 + mix)
    sp = oracle.PacketView(4)
    iq = oracle.synth_iq(fmt, 4)
    L.orc_converter_mix(c, iq, len(iq), sp.p, 100_000_000)
    n = L.orc_converter_nco_len(c)
    oc, os_ = np.empty(max(n, 1), np.float32), np.empty(max(n, 1), np.float32)
    L.orc_converter_nco_table(c, oc, os_)
    eff, length = C.c_int(), C.c_int()
    pc, ps = np.zeros(500, np.float32), np.zeros(500, np.float32)
    assert lib.rfa_nco_design(fmt, fs, mix, C.byref(eff), C.byref(length), pc.ctypes.data, ps.ctypes.data) == 0
    assert eff.value == L.orc_converter_nco_freq(c) and length.value == n
    assert np.array_equal(pc[:n], oc[:n]) and np.array_equal(ps[:n], os_[:n])
    L.orc_converter_free(c)


def test_channel_bins(lib, oracle):
    b0, b1 = C.c_int(), C.c_int()
    assert lib.rfa_channel_bins(4096, 100_000_000, 20_000_000, 100_240_000, 100_260_000, C.byref(b0), C.byref(b1)) == 0
    mag = np.arange(4096, dtype=np.float32)
    out = C.c_float()
    assert oracle.lib().orc_signal_strength(mag, 4096, 100_000_000, 20_000_000, 100_240_000, 100_260_000, C.byref(out)) == 1
    assert b1.value > b0.value
    assert abs(out.value - mag[b0.value:b1.value].mean()) < 1e-3


LOWPASS_CASES = [(1.0, 1000.0, 100.0, 50.0, 60.0), (1.0, 1000.0, 100.0, 100.0, 40.0),      # ApplicationTest.kt
                 (1.0, 384000.0, 100000.0, 38400.0, 60.0), (1.0, 96000.0, 10000.0, 9600.0, 60.0),  # user filters
                 (1.0, 1.0, 0.1, 0.15, 30.0), (1.0, 1.0, 0.1, 0.1, 30.0),                   # AudioSink.java:94-96
                 (1.0, 48000.0, 9000.0, 3000.0, 60.0)]


@pytest.mark.parametrize("case", LOWPASS_CASES)
def test_lowpass_taps_bit_exact(lib, oracle, case):
    want = oracle.lowpass_taps(*case)
    n = C.c_int()
    buf = np.zeros(4096, np.float32)
    assert lib.rfa_design_lowpass(*case, _lib.TAPWIN_BLACKMAN, 0.0, 0, buf.ctypes.data, len(buf), C.byref(n)) == 0
    assert n.value == len(want) and np.array_equal(buf[: n.value], want)


def test_lowpass_firdes_checks(lib):
    n = C.c_int()
    assert lib.rfa_design_lowpass(1.0, 1000.0, 600.0, 50.0, 60.0, 0, 0.0, 0, None, 0, C.byref(n)) == _lib.ERR_INVALID
    assert lib.rfa_design_lowpass(1.0, 1000.0, 100.0, 0.0, 60.0, 0, 0.0, 0, None, 0, C.byref(n)) == _lib.ERR_INVALID


@pytest.mark.parametrize("fs,lo,hi,tw,att", [(96000.0, 200.0, 2800.0, 960.0, 40.0), (96000.0, -2800.0, -200.0, 960.0, 40.0),
                                            (48000.0, 600.0, 900.0, 480.0, 40.0)])
def test_bandpass_taps_bit_exact(lib, oracle, fs, lo, hi, tw, att):
    wr, wi = oracle.bandpass_taps(1.0, fs, lo, hi, tw, att)
    n = C.c_int()
    tre, tim = np.zeros(4096, np.float32), np.zeros(4096, np.float32)
    assert lib.rfa_design_bandpass(1.0, fs, lo, hi, tw, att, tre.ctypes.data, tim.ctypes.data, 4096, C.byref(n)) == 0
    assert n.value == len(wr) == 181
    assert np.array_equal(tre[: n.value], wr) and np.array_equal(tim[: n.value], wi)


@pytest.mark.parametrize("i,d,maxtaps", [(4, 25, 500), (6, 625, 500), (3, 625, 500), (12, 625, 500), (11, 17, 0), (17, 11, 0)])
def test_resampler_taps_bit_exact(lib, oracle, i, d, maxtaps):
    want = oracle.resampler_taps(i, d, 0.4, maxtaps)
    n = C.c_int()
    buf = np.zeros(1 << 16, np.float32)
    assert lib.rfa_design_resampler_taps(i, d, 0.4, maxtaps, buf.ctypes.data, len(buf), C.byref(n)) == 0
    assert n.value == len(want) and np.array_equal(buf[: n.value], want)


def test_tap_windows_and_limit_denominator(lib, oracle):
    L = oracle.lib()
    out = C.c_float()
    for kind, beta in ((0, 0.0), (1, 0.0), (2, 7.0), (2, 0.0)):
        for N in (9, 55, 821):
            for n in (0, 1, N // 2, N - 1):
                assert lib.rfa_tap_window(kind, beta, n, N, C.byref(out)) == 0
                assert out.value == L.orc_window_value(kind, beta, n, N)
    a, b = C.c_int(), C.c_int()
    for num, den in ((384000, 2400000), (96000, 10000000), (2500101, 250000), (48000, 10000000), (96000, 2048000),
                     (384000, 20000000), (1234567, 7654321)):
        lib.rfa_limit_denominator(num, den, 10000, C.byref(a), C.byref(b))
        assert (a.value, b.value) == oracle.limit_denominator(num, den, 10000)


def test_mode_table(lib, oracle):
    q, lo, hi, de = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    L = oracle.lib()
    for mode in range(1, 7):
        assert lib.rfa_mode_info(mode, C.byref(q), C.byref(lo), C.byref(hi), C.byref(de)) == 0
        assert q.value == L.orc_mode_quadrature_rate(mode)
        d = L.orc_demod_new(16)
        L.orc_demod_set_mode(d, mode)
        assert de.value == L.orc_demod_channel_width(d)
        L.orc_demod_set_channel_width(d, 1)
        assert lo.value == L.orc_demod_channel_width(d)
        L.orc_demod_set_channel_width(d, 10 ** 7)
        assert hi.value == L.orc_demod_channel_width(d)
        L.orc_demod_free(d)
