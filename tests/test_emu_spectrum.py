"""CPU emulation of the fused spectrum kernel (the kernel's own __host__ __device__ phase
functions, run thread by thread) against the oracle.  Guards the FFT index logic, twiddle
tables, bit-exact conversion and dB scaling on the build box, where there is no GPU."""
import numpy as np
import pytest

SIZES = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]


def run_emu(emu, n, fmt, out_kind, window, x, frames, want_peaks=False, avg_len=None, x_im=None):
    rows = np.zeros((frames, n * (2 if out_kind == 1 else 1)), np.float32)
    peaks = np.zeros(n, np.float32) if want_peaks else None
    avg = np.zeros(n, np.float32) if avg_len is not None else None
    rc = emu.emu_spectrum_avg(n, fmt, out_kind, window, x.ctypes.data, None if x_im is None else x_im.ctypes.data,
                              frames, rows.ctypes.data, None if peaks is None else peaks.ctypes.data,
                              None if avg is None else avg.ctypes.data, avg_len or 0)
    assert rc == 0
    return rows, peaks, avg


@pytest.mark.parametrize("n", SIZES)
def test_complex_fft_all_sizes(emu, n):
    """performFFT contract (nativedsp.cpp:19-42): ordered forward C2C, unnormalised."""
    rng = np.random.default_rng(n)
    frames = 2
    x = (rng.standard_normal((frames, n)) + 1j * rng.standard_normal((frames, n))).astype(np.complex64)
    rows, _, _ = run_emu(emu, n, 3, 1, -1, x, frames)
    got = rows.view(np.complex64)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-6


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("n", [16, 256, 1024, 4096, 8192, 32768, 65536])
def test_fused_path_vs_oracle(emu, oracle, fmt, n):
    frames = 3 if n <= 8192 else 2
    iq = oracle.synth_iq(fmt, n * frames)
    rows, peaks, avg = run_emu(emu, n, fmt, 0, 0, iq, frames, want_peaks=True, avg_len=1)
    r, p, a = oracle.spectrum_run(fmt, iq, n, 1)
    assert np.abs(rows - r).max() < 0.01      # dB, north-star tolerance
    assert np.abs(peaks - p).max() < 0.01
    assert np.abs(avg - a).max() < 0.01
    lin, lin_ref = 10.0 ** (rows / 5.0), 10.0 ** (r / 5.0)     # |X|^2/N^2
    assert np.all(np.abs(lin - lin_ref) <= 1e-4 * lin_ref + 1e-6 * lin_ref.max())


def test_extreme_codes_and_zero_frame(emu, oracle):
    """Edge inputs: full-scale codes, and an all-zero s8 frame (|X| = 0 -> -inf dB like log10f(0))."""
    n = 1024
    iq = np.zeros(2 * n, np.uint8)
    rows, _, _ = run_emu(emu, n, 0, 0, 0, iq, 1)
    assert np.all(np.isneginf(rows))
    iq = np.tile(np.array([0x80, 0x7F], np.uint8), n)  # I = -128, Q = +127
    rows, _, _ = run_emu(emu, n, 0, 0, 0, iq, 1)
    r, _, _ = oracle.spectrum_run(0, iq, n, 0)
    # a windowed constant: everything but the main lobe is float32 rounding noise ~130 dB
    # down, where neither pffft builds nor this kernel agree bin by bin
    big = r > r.max() - 50
    assert big.sum() >= 5 and np.abs(rows[big] - r[big]).max() < 0.01
    lin, lin_ref = 10.0 ** (rows / 5.0), 10.0 ** (r / 5.0)
    assert np.all(np.abs(lin - lin_ref) <= 1e-4 * lin_ref + 1e-6 * lin_ref.max())


def test_planar_windowed_entry(emu, oracle):
    """NativeDsp.performWindowedFftAndReturnMag (NativeDsp.kt:43-62): planar float in, dB out."""
    n = 2048
    rng = np.random.default_rng(5)
    re = rng.standard_normal(n).astype(np.float32)
    im = rng.standard_normal(n).astype(np.float32)
    rows, _, _ = run_emu(emu, n, 4, 0, 0, re, 1, x_im=im)
    mag = np.empty(n, np.float32)
    assert oracle.lib().orc_windowed_fft_logmag(re, im, n, n, n, mag) == 1
    assert np.abs(rows[0] - mag).max() < 0.01


# ---- dual-frame kernel (spectrum2_kernel.cuh): two consecutive frames per thread, N = 256 .. 4096 ----
def run_emu2(emu, n, fmt, window, x, frames, avg_len=1, store_from=0):
    rows = np.full((frames, n), 7.0, np.float32)
    peaks = np.zeros(n, np.float32)
    avg = np.zeros(n, np.float32)
    rc = emu.emu_spectrum2_avg(n, fmt, window, x.ctypes.data, frames, rows.ctypes.data, peaks.ctypes.data,
                               avg.ctypes.data, avg_len, store_from)
    assert rc == 0
    return rows, peaks, avg


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("n,frames", [(256, 4), (512, 5), (1024, 3), (2048, 2), (4096, 3), (4096, 1)])
def test_dual_frame_path_vs_oracle(emu, oracle, fmt, n, frames):
    """Even and odd frame counts (an odd count leaves the oldest frame without a partner)."""
    iq = oracle.synth_iq(fmt, n * frames)
    rows, peaks, avg = run_emu2(emu, n, fmt, 0, iq, frames)
    r, p, a = oracle.spectrum_run(fmt, iq, n, 1)
    assert np.abs(rows - r).max() < 0.01
    assert np.abs(peaks - p).max() < 0.01
    assert np.abs(avg - a).max() < 0.01
    lin, lin_ref = 10.0 ** (rows / 5.0), 10.0 ** (r / 5.0)
    assert np.all(np.abs(lin - lin_ref) <= 1e-4 * lin_ref + 1e-6 * lin_ref.max())


def test_dual_frame_matches_single_frame_kernel(emu, oracle):
    """Both kernels run the same butterflies in the same order (only the rounding of the twiddle
    products differs): rows agree far inside the parity tolerance."""
    n, frames = 2048, 5
    iq = oracle.synth_iq(0, n * frames)
    rows1, peaks1, _ = run_emu(emu, n, 0, 0, 0, iq, frames, want_peaks=True, avg_len=1)
    rows2, peaks2, _ = run_emu2(emu, n, 0, 0, iq, frames)
    assert np.abs(rows1 - rows2).max() < 2e-4 and np.abs(peaks1 - peaks2).max() < 2e-4


def test_dual_frame_store_from(emu, oracle):
    """Rows of frames below store_from stay untouched, peaks still cover every frame."""
    n, frames = 1024, 5
    iq = oracle.synth_iq(1, n * frames)
    rows, peaks, _ = run_emu2(emu, n, 1, 0, iq, frames, store_from=2)
    r, p, _ = oracle.spectrum_run(1, iq, n, 1)
    assert np.all(rows[:2] == 7.0)
    assert np.abs(rows[2:] - r[2:]).max() < 0.01 and np.abs(peaks - p).max() < 0.01


# ---- two-pass 64 x 64 kernel (spectrum64_kernel.cuh), N = 4096 ----
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_k64_path_vs_oracle(emu, oracle, fmt):
    n, frames = 4096, 3
    iq = oracle.synth_iq(fmt, n * frames)
    rows = np.zeros((frames, n), np.float32)
    peaks = np.zeros(n, np.float32)
    assert emu.emu_spectrum64(fmt, 0, iq.ctypes.data, frames, rows.ctypes.data, peaks.ctypes.data) == 0
    r, p, _ = oracle.spectrum_run(fmt, iq, n, 0)
    assert np.abs(rows - r).max() < 0.01
    assert np.abs(peaks - p).max() < 0.01
    lin, lin_ref = 10.0 ** (rows / 5.0), 10.0 ** (r / 5.0)
    assert np.all(np.abs(lin - lin_ref) <= 1e-4 * lin_ref + 1e-6 * lin_ref.max())


# ---- four-step path (fourstep_kernel.cuh): N = 32768 / 65536 as N1 x 256 through an intermediate buffer ----
@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("n", [32768, 65536])
def test_fourstep_vs_oracle(emu, oracle, fmt, n):
    frames = 2
    iq = oracle.synth_iq(fmt, n * frames)
    rows = np.full((frames, n), 7.0, np.float32)
    peaks = np.zeros(n, np.float32)
    assert emu.emu_fourstep_spectrum(n, fmt, 0, iq.ctypes.data, frames, rows.ctypes.data, peaks.ctypes.data, 0) == 0
    r, p, _ = oracle.spectrum_run(fmt, iq, n, 1)
    assert np.abs(rows - r).max() < 0.01
    assert np.abs(peaks - p).max() < 0.01
    lin, lin_ref = 10.0 ** (rows / 5.0), 10.0 ** (r / 5.0)
    assert np.all(np.abs(lin - lin_ref) <= 1e-4 * lin_ref + 1e-6 * lin_ref.max())


def test_fourstep_store_from_keeps_older_rows_untouched(emu, oracle):
    n, frames = 32768, 3
    iq = oracle.synth_iq(0, n * frames)
    rows = np.full((frames, n), 7.0, np.float32)
    peaks = np.zeros(n, np.float32)
    assert emu.emu_fourstep_spectrum(n, 0, 0, iq.ctypes.data, frames, rows.ctypes.data, peaks.ctypes.data, 2) == 0
    r, p, _ = oracle.spectrum_run(0, iq, n, 1)
    assert np.all(rows[:2] == 7.0)
    assert np.abs(rows[2] - r[2]).max() < 0.01
    assert np.abs(peaks - p).max() < 0.01      # peak hold still sees every frame


# ---- cluster path (fourstep_cluster.cuh): the same factorisation, the intermediate in the CTAs' shared memory ----
@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("n", [32768, 65536])
def test_cluster_path_vs_oracle_and_two_kernel_path(emu, oracle, fmt, n):
    """Ownership of columns / rows by cluster rank, the dense step-A exchange and the XOR-swizzled step-B exchange,
    thread by thread; the butterflies are those of the two-kernel path, so the rows are identical bit for bit."""
    frames = 2
    iq = oracle.synth_iq(fmt, n * frames)
    rows = np.full((frames, n), 7.0, np.float32)
    peaks = np.zeros(n, np.float32)
    assert emu.emu_cluster_spectrum(n, fmt, 0, iq.ctypes.data, frames, rows.ctypes.data, peaks.ctypes.data, 0) == 0
    r, p, _ = oracle.spectrum_run(fmt, iq, n, 1)
    assert np.abs(rows - r).max() < 0.01
    assert np.abs(peaks - p).max() < 0.01
    lin, lin_ref = 10.0 ** (rows / 5.0), 10.0 ** (r / 5.0)
    assert np.all(np.abs(lin - lin_ref) <= 1e-4 * lin_ref + 1e-6 * lin_ref.max())
    rows2 = np.zeros((frames, n), np.float32)
    peaks2 = np.zeros(n, np.float32)
    assert emu.emu_fourstep_spectrum(n, fmt, 0, iq.ctypes.data, frames, rows2.ctypes.data, peaks2.ctypes.data, 0) == 0
    assert np.array_equal(rows, rows2) and np.array_equal(peaks, peaks2)


def test_cluster_path_store_from(emu, oracle):
    n, frames = 65536, 3
    iq = oracle.synth_iq(2, n * frames)
    rows = np.full((frames, n), 7.0, np.float32)
    peaks = np.zeros(n, np.float32)
    assert emu.emu_cluster_spectrum(n, 2, 0, iq.ctypes.data, frames, rows.ctypes.data, peaks.ctypes.data, 2) == 0
    r, p, _ = oracle.spectrum_run(2, iq, n, 1)
    assert np.all(rows[:2] == 7.0)
    assert np.abs(rows[2] - r[2]).max() < 0.01
    assert np.abs(peaks - p).max() < 0.01
