"""dsp.FftProcessor (the GPU-side mirror of FftProcessor.run, FftProcessor.kt:111-253) against the
oracle's orc_fftproc: ring maintenance across retunes (history shift, :199-217), sample-rate change
(:218-222), peak-hold reset (:236-241) and the channel signal strength (:143-157)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ring_shift_on_retune_and_signal_strength(gpu_ctx, oracle):
    import torch
    import rfanalyzer_b200 as rfa
    L = oracle.lib()
    n, fmt = 1024, 0
    strengths = []
    with torch.cuda.stream(gpu_ctx.torch_stream):
        data = rfa.FftProcessorData()
        proc = rfa.FftProcessor(gpu_ctx, data, waterfallSpeed="FAST", fftPeakHold=True,
                                getChannelFrequencyRange=lambda: (100_020_000, 100_060_000),
                                onAverageSignalStrengthChanged=strengths.append)
        ref = L.orc_fftproc_new(300, 1)
        first = 0
        # (frames, frequency, sample rate): steady, retune up (shift left), retune down (shift right),
        # retune far away (history cleared), sample-rate change (history cleared)
        plan = [(5, 100_000_000, 1_000_000), (3, 100_050_000, 1_000_000), (2, 99_990_000, 1_000_000),
                (2, 150_000_000, 1_000_000), (4, 150_000_000, 2_000_000)]
        for frames, freq, fs in plan:
            iq = oracle.synth_iq(fmt, n * frames, first=first)
            first += n * frames
            rows, _, _ = oracle.spectrum_run(fmt, iq, n, 0)
            want_strength = None
            for k in range(frames):
                L.orc_fftproc_push(ref, np.ascontiguousarray(rows[k]), n, freq, fs)
            import ctypes as C
            out = C.c_float()
            if L.orc_signal_strength(np.ascontiguousarray(rows[frames - 1]), n, freq, fs, 100_020_000, 100_060_000, C.byref(out)):
                want_strength = out.value
            got_before = len(strengths)
            # the reference pushes frame by frame; a retune happens between packets, so one
            # batched call per (frequency, rate) segment is the same sequence of ring updates
            proc.process_iq(fmt, torch.from_numpy(iq).cuda(), frames, n, freq, fs)
            gpu_ctx.sync()
            ring_ref = np.stack([np.ctypeslib.as_array(L.orc_fftproc_row(ref, i), shape=(n,)) for i in range(300)])
            got = data.waterfallBuffer.cpu().numpy()
            assert np.abs(got - ring_ref).max() < 0.01, (freq, fs)
            assert data.writeIndex == L.orc_fftproc_write_index(ref) and data.readIndex == L.orc_fftproc_read_index(ref)
            peaks_ref = np.ctypeslib.as_array(L.orc_fftproc_peaks(ref), shape=(n,))
            assert np.abs(data.peaks.cpu().numpy() - peaks_ref).max() < 0.01
            if want_strength is not None:
                assert len(strengths) == got_before + 1 and abs(strengths[-1] - want_strength) < 0.01
        L.orc_fftproc_free(ref)


def test_shift_rows_kernel(gpu_ctx):
    """System.arraycopy + fill(-9999f) of FftProcessor.kt:203-210, both directions, in place."""
    import torch

    def ref_shift(a, s):
        out = np.full_like(a, -9999.0)
        if s >= 0:
            out[:, s:] = a[:, : a.shape[1] - s]
        else:
            out[:, : a.shape[1] + s] = a[:, -s:]
        return out

    with torch.cuda.stream(gpu_ctx.torch_stream):
        rows = torch.arange(3 * 64, dtype=torch.float32, device="cuda").reshape(3, 64).contiguous()
        want = rows.cpu().numpy()
        for s in (5, -7, 0, 63, -64):
            gpu_ctx.shift_rows(rows, 3, 64, 64, s)
            gpu_ctx.sync()
            want = ref_shift(want, s)
            assert np.array_equal(rows.cpu().numpy(), want)
