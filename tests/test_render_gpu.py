"""rfa_render_waterfall against the oracle's restatement of
AnalyzerSurface.drawPreprocessing (AnalyzerSurface.kt:646-734): colour indices, FFT trace and peak trace,
bit for bit, for full, zoomed and shifted viewports, host and device outputs."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fill_ring(oracle, rfa, gpu_ctx, n, ring, frames, fs, freq):
    import torch
    L_o = oracle.lib()
    proc = L_o.orc_fftproc_new(ring, 1)
    iq = oracle.synth_iq(0, n * frames)
    r, _, _ = oracle.spectrum_run(0, iq, n, 0)
    for k in range(frames):
        L_o.orc_fftproc_push(proc, np.ascontiguousarray(r[k]), n, freq, fs)
    rows = np.stack([np.ctypeslib.as_array(L_o.orc_fftproc_row(proc, i), shape=(n,)) for i in range(ring)]).copy()
    peaks = np.ctypeslib.as_array(L_o.orc_fftproc_peaks(proc), shape=(n,)).copy()
    return proc, rows, peaks, L_o.orc_fftproc_read_index(proc)


@pytest.mark.parametrize("vp_freq_off,vp_rate,width", [
    (0, 20_000_000, 1080),           # whole span, ~3.8 bins per pixel
    (0, 20_000_000, 4096),           # one bin per pixel
    (1_500_000, 5_000_000, 1080),    # zoomed in and shifted: about one bin per pixel
    (-6_000_000, 20_000_000, 777),   # shifted left: black pixels on the left (firstPixel > 0)
    (7_000_000, 30_000_000, 1920),   # zoomed out and shifted: black on both sides
])
@pytest.mark.parametrize("host_out", [True, False])
def test_render_matches_the_reference_arithmetic(gpu_ctx, oracle, vp_freq_off, vp_rate, width, host_out):
    import torch
    import rfanalyzer_b200 as rfa
    n, ring, frames, fs, freq, L, height, cmsize = 4096, 300, 37, 20_000_000, 100_000_000, 5, 600, 256
    min_db, max_db = -52.5, -3.25
    proc, rows, peaks, read_index = _fill_ring(oracle, rfa, gpu_ctx, n, ring, frames, fs, freq)
    L_o = oracle.lib()
    want_avg = np.zeros(width, np.float32)
    want_idx = np.zeros((ring, width), np.int32)
    want_py = np.zeros(width, np.float32)
    L_o.orc_draw_preprocess(proc, width, height, freq + vp_freq_off, vp_rate, min_db, max_db, L, cmsize,
                            want_avg, want_idx, want_py.ctypes.data)
    cmap = (np.arange(cmsize, dtype=np.uint32) * 0x010203 + 0xFF000000).astype(np.uint32)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d_rows = torch.from_numpy(rows).cuda()
        d_peaks = torch.from_numpy(peaks).cuda()
        if host_out:
            argb = np.zeros((ring, width), np.uint32)
            idx = np.zeros((ring, width), np.int32)
            tavg, py = np.zeros(width, np.float32), np.zeros(width, np.float32)
            cm = cmap
        else:
            argb = torch.zeros((ring, width), dtype=torch.int32, device="cuda")
            idx = torch.zeros((ring, width), dtype=torch.int32, device="cuda")
            tavg = torch.zeros(width, dtype=torch.float32, device="cuda")
            py = torch.zeros(width, dtype=torch.float32, device="cuda")
            cm = torch.from_numpy(cmap.view(np.int32)).cuda()
        gpu_ctx.render_waterfall(d_rows, d_peaks, n, freq, fs, freq + vp_freq_off, vp_rate, width, height, min_db,
                                 max_db, L, ring, read_index, colormap=cm, colormap_size=cmsize, argb=argb,
                                 color_index=idx, time_average=tavg, peaks_y=py)
        gpu_ctx.sync()
    if not host_out:
        argb, idx, tavg, py = (argb.cpu().numpy().view(np.uint32), idx.cpu().numpy(), tavg.cpu().numpy(), py.cpu().numpy())
    assert np.array_equal(idx, want_idx)
    inside = want_idx >= 0
    assert inside.any()
    assert np.array_equal(argb[inside], cmap[want_idx[inside]]) and np.all(argb[~inside] == 0xFF000000)
    drawn = ~np.isnan(want_avg)
    assert np.array_equal(np.isnan(tavg), ~drawn) and np.array_equal(tavg[drawn], want_avg[drawn])
    assert np.array_equal(py, want_py)
    L_o.orc_fftproc_free(proc)


def test_render_dirty_rows_only(gpu_ctx, oracle):
    """Rendering a row range leaves the other rows of the colour buffer untouched (AnalyzerSurface.kt:688-691)."""
    import torch
    import rfanalyzer_b200 as rfa
    n, ring, frames, fs, freq = 1024, 300, 20, 2_000_000, 433_000_000
    proc, rows, peaks, read_index = _fill_ring(oracle, rfa, gpu_ctx, n, ring, frames, fs, freq)
    width = 500
    idx = np.full((ring, width), 12345, np.int32)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        gpu_ctx.render_waterfall(torch.from_numpy(rows).cuda(), None, n, freq, fs, freq, fs, width, 400, -60.0, -10.0, 0,
                                 ring, read_index, colormap_size=200, color_index=idx, first_row=2, nrows=3)
        gpu_ctx.sync()
    touched = {(read_index + k) % ring for k in (2, 3, 4)}
    for r in range(ring):
        assert (not np.all(idx[r] == 12345)) == (r in touched)
    oracle.lib().orc_fftproc_free(proc)
