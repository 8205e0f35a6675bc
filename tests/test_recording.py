"""Recordings on disk: file-name metadata and FileIQSource packet semantics against the
oracle's restatement (oracle/recording.py, the reference's own regular expressions); no GPU needed -- these are
host functions of librfa_b200.so."""
import os

import numpy as np
import pytest

from oracle import recording as R
from rfanalyzer_b200 import _lib
from rfanalyzer_b200.dsp import FileIQSource, parse_recording_name, recording_file_name

NAMES = [
    "20250111-143022_MyRecording_AIRSPY_100MHz_6MSps.iq",      # IQ_FILE_FORMAT.md:99
    "20250101_test_HACKRF_100MHz_2Msps.iq",                    # SourceTab.kt:793
    "ongoing_recording.iq", "recording.iq",                    # nothing to extract
    "20240229-235959_a b_RTLSDR_433920kHz_2400kSps.iq",
    "capture_rtl-sdr_868300000Hz_1024000sps.iq",
    "x_HydraSDR-10MSPS-1090MHZ.iq",
    "hackrfone 2400 MHz_20MSps.iq",                            # digits not directly before the unit
    "gqrx_20200101_145000000_2000000_fc.raw",
    "foo_AIRSPY_hackrf_2.4GHz_6MSps.iq",                       # later format wins; GHz is not recognised
    "a_1Hz_2Hz_3kHz.iq", "a_12MHz-7MHz.iq", "_5Sps_6ksps", "test-100khz-250KSPS_RTLSDR",
    "n_99999999999999999999Hz_5MSps_hackrf.iq",                # NumberFormatException: later fields stay unset
    "weird_airspy\nname_10MSps_100MHz.iq",
]


def lib():
    return _lib.load()


@pytest.mark.parametrize("name", NAMES)
def test_parse_name_matches_the_reference_regexes(name):
    for defaults in ((R.HACKRF, 97_000_000, 1_000_000), (R.AIRSPY, 0, 0)):
        assert parse_recording_name(lib(), name, *defaults) == R.parse_name(name, *defaults)


@pytest.mark.parametrize("freq,rate", [(100_000_000, 6_000_000), (433_920_000, 2_400_000), (2_400_000_000, 20_000_000),
                                       (1_090_000_001, 10_000_000), (0, 1), (999, 1000), (5_000_000_000_000_000, 2_500_000)])
def test_file_name_round_trip(freq, rate):
    for ff in range(4):
        name = recording_file_name(lib(), "20250111-143022", "My Rec", ff, freq, rate)
        assert name == R.calculate_file_name("20250111-143022", "My Rec", ff, freq, rate)
        got = parse_recording_name(lib(), name, R.HACKRF, -1, -1)
        assert got == R.parse_name(name, R.HACKRF, -1, -1)
        assert got[0] == ff
    assert R.calculate_file_name("20250111-143022", "MyRecording", R.AIRSPY, 100_000_000, 6_000_000) == NAMES[0]


def test_sample_formats():
    L = lib()
    assert [L.rfa_recording_sample_format(f) for f in range(5)] == [0, 1, 2, 2, -1]


@pytest.mark.parametrize("size,packet,repeat", [(10_000, 4096, False), (10_000, 4096, True), (8192, 4096, True),
                                                (100, 4096, True), (0, 16, False), (4096, 4096, False)])
def test_file_source_packets(tmp_path, size, packet, repeat):
    """getPacket: whole packets only (the trailing partial packet is dropped), rewind on repeat."""
    data = np.random.default_rng(size).integers(0, 256, size, dtype=np.uint8)
    path = os.path.join(tmp_path, "x_RTLSDR_100MHz_2MSps.iq")
    data.tofile(path)
    src = FileIQSource(lib(), path, 2_000_000, 100_000_000, packet, repeat, FileIQSource.FILE_FORMAT_8BIT_UNSIGNED)
    assert src.getBytesPerSample() == 2 and src.getPacketSize() == packet
    want = R.packets(data.tobytes(), packet, repeat, 7)
    for w in want:
        got = src.getPacket()
        assert (got is None) == (w is None)
        if w is not None:
            assert got.tobytes() == w
    src.close()


def test_file_source_paces_like_the_hardware(tmp_path):
    import time
    path = os.path.join(tmp_path, "p.iq")
    np.zeros(40_000, np.uint8).tofile(path)
    src = FileIQSource(lib(), path, 100_000, 0, 4000, False, FileIQSource.FILE_FORMAT_8BIT_SIGNED, pace=True)
    t0 = time.perf_counter()
    n = 0
    while src.getPacket() is not None:
        n += 1
    dt = time.perf_counter() - t0
    assert n == 10 and 0.18 <= dt < 1.0     # 20 000 samples at 100 kS/s = 0.2 s (FileIQSource.java:343-347)


def test_file_source_errors(tmp_path):
    with pytest.raises(_lib.RfaError):
        FileIQSource(lib(), os.path.join(tmp_path, "missing.iq"), 1, 0, 16, False, 0)
