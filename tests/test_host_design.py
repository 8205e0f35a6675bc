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
