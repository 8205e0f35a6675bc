"""Batch objects over the C ABI: Context (device + stream) and SpectrumPlan (fused
convert -> window -> FFT -> dB -> rows / peak hold / time average).

Buffers may be torch CUDA tensors (device mode, asynchronous on the context's stream) or
numpy arrays / CPU tensors (host mode, the library stages the copies).  One call uses one
memory space for all its buffers.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


def _is_device(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _mem_of(*bufs):
    kinds = {_is_device(b) for b in bufs if b is not None}
    if len(kinds) > 1:
        raise ValueError("all buffers of one call must live in the same memory space")
    return _lib.MEM_DEVICE if kinds == {True} else _lib.MEM_HOST


class Context:
    """A GPU plus the CUDA stream every call of this context is ordered on."""

    def __init__(self, device=0, stream=None):
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        if stream is not None and hasattr(stream, "cuda_stream"):
            # torch's default stream has handle 0 = the legacy default stream, which the C ABI
            # spells cudaStreamLegacy ((cudaStream_t)1); NULL would mean "create a private stream"
            stream = stream.cuda_stream or 1
        check(self.lib.rfa_ctx_create(int(device), stream, C.byref(self.handle)))
        self.device = int(device)

    def close(self):
        if self.handle:
            self.lib.rfa_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.rfa_ctx_sync(self.handle))

    def set_option(self, name, value):
        """A knob of this context (rfa_ctx_set_option; csrc/tuning.h).  Returns the previous value."""
        old = self.get_option(name)
        check(self.lib.rfa_ctx_set_option(self.handle, name.encode(), int(value)))
        return old

    def get_option(self, name):
        v = C.c_longlong()
        check(self.lib.rfa_ctx_get_option(self.handle, name.encode(), C.byref(v)))
        return v.value

    def options(self, **kw):
        """Context manager: set knobs for a block, restore them afterwards."""
        ctx = self

        class _Scope:
            def __enter__(self):
                self.old = {k: ctx.set_option(k, v) for k, v in kw.items()}

            def __exit__(self, *exc):
                for k, v in self.old.items():
                    ctx.set_option(k, v)
        return _Scope()

    @property
    def sm_count(self):
        return self.lib.rfa_ctx_sm_count(self.handle)

    @property
    def launch_count(self):
        return self.lib.rfa_ctx_launch_count(self.handle)

    # --- IQConverter.fillPacketIntoSamplePacket / mixPacketIntoSamplePacket -------------
    def convert(self, fmt, iq, nsamples, re, im):
        check(self.lib.rfa_convert(self.handle, fmt, ptr(iq), int(nsamples), ptr(re), ptr(im), _mem_of(iq, re, im)))

    def nco_design(self, fmt, sample_rate, mix_frequency):
        """generateMixerLookupTable: -> (effective_frequency, cos[len], sin[len])."""
        eff, length = C.c_int(), C.c_int()
        cos_t = np.zeros(500, np.float32)
        sin_t = np.zeros(500, np.float32)
        check(self.lib.rfa_nco_design(fmt, int(sample_rate), int(mix_frequency), C.byref(eff), C.byref(length),
                                      ptr(cos_t), ptr(sin_t)))
        return eff.value, cos_t[: length.value].copy(), sin_t[: length.value].copy()

    def mix(self, fmt, iq, nsamples, cos_t, sin_t, nco_index, re, im):
        cos_t = np.ascontiguousarray(cos_t, np.float32)
        sin_t = np.ascontiguousarray(sin_t, np.float32)
        check(self.lib.rfa_mix(self.handle, fmt, ptr(iq), int(nsamples), ptr(cos_t), ptr(sin_t), len(cos_t),
                               int(nco_index), ptr(re), ptr(im), _mem_of(iq, re, im)))

    # --- NativeDsp ------------------------------------------------------------------------
    def make_window(self, window, n):
        w = np.empty(n, np.float32)
        check(self.lib.rfa_make_window(window, n, ptr(w)))
        return w

    def fft_c2c(self, x, out, n, batch=1):
        check(self.lib.rfa_fft_c2c(self.handle, ptr(x), ptr(out), int(n), int(batch), _mem_of(x, out)))

    def fft_logmag(self, x, mag, n, batch=1):
        check(self.lib.rfa_fft_logmag(self.handle, ptr(x), ptr(mag), int(n), int(batch), _mem_of(x, mag)))

    def windowed_fft_logmag(self, re, im, mag, n, batch=1, window=_lib.WIN_BLACKMAN_REF):
        check(self.lib.rfa_windowed_fft_logmag(self.handle, ptr(re), ptr(im), ptr(mag), int(n), int(batch),
                                               int(window), _mem_of(re, im, mag)))

    # --- row reductions --------------------------------------------------------------------
    def average_rows(self, rows, newest, direction, ring_rows, row_stride, valid, avg_len, n, avg):
        check(self.lib.rfa_average_rows(self.handle, ptr(rows), int(newest), int(direction), int(ring_rows),
                                        int(row_stride), int(valid), int(avg_len), int(n), ptr(avg),
                                        _lib.MEM_DEVICE, _mem_of(avg)))

    def ema_rows(self, rows, row0, row_step, ring_rows, row_stride, first, last, alpha, from_state, n, avg):
        check(self.lib.rfa_ema_rows(self.handle, ptr(rows), int(row0), int(row_step), int(ring_rows), int(row_stride),
                                    int(first), int(last), float(alpha), 1 if from_state else 0, int(n), ptr(avg),
                                    _mem_of(avg)))

    def channel_bins(self, n, frequency, sample_rate, chan_start, chan_end):
        b0, b1 = C.c_int(), C.c_int()
        check(self.lib.rfa_channel_bins(int(n), int(frequency), int(sample_rate), int(chan_start), int(chan_end),
                                        C.byref(b0), C.byref(b1)))
        return b0.value, b1.value

    def channel_strength(self, rows, row0, row_step, ring_rows, row_stride, nrows, b0, b1, out):
        check(self.lib.rfa_channel_strength(self.handle, ptr(rows), int(row0), int(row_step), int(ring_rows),
                                            int(row_stride), int(nrows), int(b0), int(b1), ptr(out), _mem_of(out)))

    def detect_windows(self, rows, row_stride, n, windows, peak=None, avg=None):
        """(peak, avg) of every (row, start, end) window over device rows (rfa_detect_windows).  `windows` is a
        sequence of triples or a ctypes array of DetectWindow; outputs default to new host arrays."""
        if not isinstance(windows, C.Array):
            windows = (_lib.DetectWindow * len(windows))(*[_lib.DetectWindow(int(r), int(a), int(b)) for r, a, b in windows])
        nwin = len(windows)
        if peak is None:
            peak, avg = np.empty(nwin, np.float32), np.empty(nwin, np.float32)
        check(self.lib.rfa_detect_windows(self.handle, ptr(rows), int(row_stride), int(n), C.addressof(windows), nwin,
                                          ptr(peak), ptr(avg), _lib.MEM_HOST, _mem_of(peak, avg)))
        return peak, avg

    def shift_rows(self, rows, nrows, row_stride, n, shift):
        check(self.lib.rfa_shift_rows(self.handle, ptr(rows), int(nrows), int(row_stride), int(n), int(shift)))

    def render_waterfall(self, rows, peaks, n, frequency, sample_rate, viewport_frequency, viewport_sample_rate,
                         width, fft_height, min_db, max_db, avg_len, ring_rows, newest_row, colormap=None,
                         colormap_size=None, argb=None, color_index=None, time_average=None, peaks_y=None,
                         first_row=0, nrows=None, row_stride=0):
        """AnalyzerSurface.drawPreprocessing on the device ring `rows` (rfa_render_waterfall).  Outputs are
        host arrays or device tensors, all of one kind."""
        d = _lib.RenderDesc(int(n), int(frequency), int(sample_rate), int(viewport_frequency),
                            int(viewport_sample_rate), int(width), int(fft_height), float(min_db), float(max_db),
                            int(avg_len), int(ring_rows), int(row_stride), int(newest_row), int(first_row),
                            int(ring_rows - first_row if nrows is None else nrows))
        outs = [o for o in (argb, color_index, time_average, peaks_y) if o is not None]
        size = colormap_size if colormap_size is not None else (len(colormap) if colormap is not None else 0)
        check(self.lib.rfa_render_waterfall(self.handle, C.byref(d), ptr(rows), ptr(peaks), ptr(colormap), int(size),
                                            _mem_of(colormap) if colormap is not None else _lib.MEM_HOST,
                                            ptr(argb), ptr(color_index), ptr(time_average), ptr(peaks_y),
                                            _mem_of(*outs) if outs else _lib.MEM_HOST))

    def fill(self, dst, count, value):
        check(self.lib.rfa_fill(self.handle, ptr(dst), int(count), float(value)))


SYNTH_SEED = 0x52464131


def synth_step(cycles_per_sample):
    """Tone frequency as a 32-bit phase increment (cycles/sample, negative allowed)."""
    frac = cycles_per_sample - np.floor(cycles_per_sample)
    return int(round(frac * 4294967296.0)) & 0xFFFFFFFF


def default_synth_components(fmt):
    """The three-tone spectrum-path signal of SURVEY.md 8(d): (f/fs, amplitude) =
    (+0.1234, 48), (-0.3071, 24), (+0.0127, 12) LSB, x256 for 16-bit samples."""
    mul = 256 if fmt == _lib.FMT_S16LE else 1
    return [(synth_step(0.1234), 48 * mul, 0, 0), (synth_step(-0.3071), 24 * mul, 0, 0),
            (synth_step(0.0127), 12 * mul, 0, 0)]


def synth_iq(ctx, fmt, nsamples, out, first=0, comps=None, noise_shift=2, seed=SYNTH_SEED):
    """Fill `out` (device tensor or host array of bytes) with samples first .. first+nsamples-1."""
    comps = default_synth_components(fmt) if comps is None else comps
    arr = (_lib.SynthComp * max(len(comps), 1))()
    for i, c in enumerate(comps):
        arr[i] = _lib.SynthComp(*[int(v) for v in c])
    check(ctx.lib.rfa_synth_iq(ctx.handle, fmt, seed, arr, len(comps), int(noise_shift), int(first), int(nsamples),
                               ptr(out), _mem_of(out)))


class SpectrumPlan:
    """Fused IQ bytes -> dB waterfall rows (+ peak hold + time average) for one FFT size."""

    def __init__(self, ctx, fmt, fft_size, window=_lib.WIN_BLACKMAN_REF, avg_len=0, peak_hold=True, ema_alpha=None):
        """ema_alpha: None = the reference's box-car mean of the newest avg_len+1 rows; a float in (0, 1] = the
        exponential average a += alpha*(row - a) over the frames in time order (RFA_AVG_EMA)."""
        self.ctx = ctx
        self.fmt, self.fft_size, self.avg_len = fmt, int(fft_size), int(avg_len)
        self.desc = _lib.SpectrumDesc(fmt, int(fft_size), int(window), int(avg_len), 1 if peak_hold else 0,
                                      _lib.AVG_BOXCAR if ema_alpha is None else _lib.AVG_EMA,
                                      0.0 if ema_alpha is None else float(ema_alpha))
        self.handle = C.c_void_p()
        check(ctx.lib.rfa_spectrum_plan_create(ctx.handle, C.byref(self.desc), C.byref(self.handle)))

    def close(self):
        if self.handle:
            self.ctx.lib.rfa_spectrum_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def algorithmic_bytes(self, nframes, rows_stored=True):
        return self.ctx.lib.rfa_spectrum_algorithmic_bytes(self.handle, int(nframes), 1 if rows_stored else 0)

    def process(self, iq, nframes, rows=None, peaks=None, avg=None, row0=0, row_step=1, ring_rows=0,
                row_stride=None, history_rows=0, peaks_accumulate=False, avg_accumulate=False):
        out = _lib.SpectrumOut(ptr(rows), int(row0), int(row_step), int(ring_rows),
                               int(row_stride if row_stride is not None else self.fft_size), int(history_rows),
                               ptr(peaks), 1 if peaks_accumulate else 0, ptr(avg), 1 if avg_accumulate else 0)
        check(self.ctx.lib.rfa_spectrum_process(self.handle, ptr(iq), int(nframes), C.byref(out),
                                                _mem_of(iq, rows, peaks, avg)))

    def process_file(self, path, first_frame=0, nframes=-1, rows=None, peaks=None, avg=None, peaks_accumulate=False,
                     chunk_frames=0):
        """Spectrum pass over a recording on disk (rfa_spectrum_process_file): host arrays for rows (all rows
        of the range, or None), peaks and avg.  Returns the number of frames transformed."""
        out = _lib.SpectrumOut(ptr(rows), 0, 1, 0, self.fft_size, 0, ptr(peaks), 1 if peaks_accumulate else 0, ptr(avg), 0)
        done = C.c_longlong()
        check(self.ctx.lib.rfa_spectrum_process_file(self.handle, path.encode(), int(first_frame), int(nframes),
                                                     C.byref(out), int(chunk_frames), C.byref(done)))
        return done.value
