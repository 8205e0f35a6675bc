"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module; the product package rfanalyzer_b200 never does.

`lib()` is the C restatement (oracle/liboracle.so, always buildable with gcc);
`ref()` is the reference's own pffft.c + nativedsp.cpp compiled in place
(oracle/_ref/librfa_ref.so, built here where /root/reference exists; the .so travels).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FMT_S8, FMT_U8, FMT_S16LE = 0, 1, 2
MODE_OFF, MODE_AM, MODE_NFM, MODE_WFM, MODE_LSB, MODE_USB, MODE_CW = range(7)
WIN_BLACKMAN, WIN_HAMMING, WIN_KAISER = 0, 1, 2
BYTES_PER_SAMPLE = {FMT_S8: 2, FMT_U8: 2, FMT_S16LE: 4}
SEED = 0x52464131

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Packet(C.Structure):
    _fields_ = [("re", C.POINTER(C.c_float)), ("im", C.POINTER(C.c_float)), ("capacity", C.c_int),
                ("size", C.c_int), ("sampleRate", C.c_int), ("frequency", C.c_longlong)]


class SynthComp(C.Structure):
    _fields_ = [("step", C.c_uint32), ("amp", C.c_int32), ("modStep", C.c_uint32), ("modK", C.c_int32)]


def build(force=False):
    """Compile the restatement and, where /root/reference exists, oracle/_ref."""
    so = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.exists("/root/reference"):
        subprocess.run(["make", "-s", "-C", HERE], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _declare(L):
    PP = C.POINTER(Packet)
    vp = C.c_void_p
    sig = {
        "orc_packet_new": (PP, [C.c_int]), "orc_packet_free": (None, [PP]),
        "orc_converter_new": (vp, [C.c_int]), "orc_converter_free": (None, [vp]),
        "orc_converter_set_frequency": (None, [vp, C.c_longlong]),
        "orc_converter_set_sample_rate": (None, [vp, C.c_int]),
        "orc_converter_lut": (C.POINTER(C.c_float), [vp, C.POINTER(C.c_int)]),
        "orc_converter_fill": (C.c_int, [vp, _u8p, C.c_int, PP]),
        "orc_converter_mix": (C.c_int, [vp, _u8p, C.c_int, PP, C.c_longlong]),
        "orc_converter_nco_len": (C.c_int, [vp]), "orc_converter_nco_index": (C.c_int, [vp]),
        "orc_converter_nco_freq": (C.c_int, [vp]),
        "orc_converter_nco_table": (None, [vp, _f32p, _f32p]),
        "orc_calc_optimal_cosine_length": (C.c_int, [C.c_int, C.c_int]),
        "orc_nativedsp_window": (None, [C.c_int, _f32p]),
        "orc_fft_c2c_f32": (None, [_f32p, _f32p, C.c_int]),
        "orc_fft_c2c_f64": (None, [_f32p, _f64p, C.c_int]),
        "orc_fft_logmag": (None, [_f32p, _f32p, C.c_int]),
        "orc_windowed_fft_logmag": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]),
        "orc_fftproc_new": (vp, [C.c_int, C.c_int]), "orc_fftproc_free": (None, [vp]),
        "orc_fftproc_push": (C.c_int, [vp, _f32p, C.c_int, C.c_longlong, C.c_int]),
        "orc_fftproc_row": (C.POINTER(C.c_float), [vp, C.c_int]),
        "orc_fftproc_peaks": (C.POINTER(C.c_float), [vp]),
        "orc_fftproc_read_index": (C.c_int, [vp]), "orc_fftproc_write_index": (C.c_int, [vp]),
        "orc_fftproc_rows": (C.c_int, [vp]),
        "orc_signal_strength": (C.c_int, [_f32p, C.c_int, C.c_longlong, C.c_int, C.c_longlong, C.c_longlong,
                                          C.POINTER(C.c_float)]),
        "orc_time_average": (None, [vp, C.c_int, _f32p]),
        "orc_ema_rows": (None, [_f32p, C.c_longlong, C.c_int, C.c_float, C.c_void_p, _f32p]),
        "orc_draw_preprocess": (None, [vp, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_float, C.c_float,
                                       C.c_int, C.c_int, _f32p, _i32p, C.c_void_p]),
        "orc_window_value": (C.c_float, [C.c_int, C.c_double, C.c_int, C.c_int]),
        "orc_lowpass_taps": (C.c_int, [C.c_float] * 5 + [C.c_int, C.c_double, C.c_int,
                                                        C.POINTER(C.POINTER(C.c_float))]),
        "orc_bandpass_taps": (C.c_int, [C.c_float] * 6 + [C.POINTER(C.POINTER(C.c_float))] * 2),
        "orc_fir_new": (vp, [_f32p, C.c_int, C.c_int]),
        "orc_fir_lowpass": (vp, [C.c_int] + [C.c_float] * 5), "orc_fir_free": (None, [vp]),
        "orc_fir_ntaps": (C.c_int, [vp]), "orc_fir_taps": (C.POINTER(C.c_float), [vp]),
        "orc_fir_filter": (C.c_int, [vp, PP, PP, C.c_int, C.c_int]),
        "orc_fir_filter_real": (C.c_int, [vp, PP, PP, C.c_int, C.c_int]),
        "orc_cfir_bandpass": (vp, [C.c_int] + [C.c_float] * 6), "orc_cfir_free": (None, [vp]),
        "orc_cfir_ntaps": (C.c_int, [vp]),
        "orc_cfir_taps_re": (C.POINTER(C.c_float), [vp]), "orc_cfir_taps_im": (C.POINTER(C.c_float), [vp]),
        "orc_cfir_filter": (C.c_int, [vp, PP, PP, C.c_int, C.c_int]),
        "orc_gcd": (C.c_int, [C.c_int, C.c_int]),
        "orc_limit_denominator": (None, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "orc_design_resampler_taps": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_int,
                                                C.POINTER(C.POINTER(C.c_float))]),
        "orc_resampler_new": (vp, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int]),
        "orc_resampler_free": (None, [vp]),
        "orc_resampler_interp": (C.c_int, [vp]), "orc_resampler_decim": (C.c_int, [vp]),
        "orc_resampler_taps_per_phase": (C.c_int, [vp]),
        "orc_resampler_bank": (None, [vp, _f32p]),
        "orc_resampler_resample": (C.c_int, [vp, PP, PP, C.c_int, C.c_int]),
        "orc_demod_new": (vp, [C.c_int]), "orc_demod_free": (None, [vp]),
        "orc_demod_set_mode": (None, [vp, C.c_int]),
        "orc_demod_set_channel_width": (None, [vp, C.c_int]),
        "orc_demod_channel_width": (C.c_int, [vp]),
        "orc_demod_set_volume": (None, [vp, C.c_float]),
        "orc_mode_quadrature_rate": (C.c_int, [C.c_int]),
        "orc_demod_process": (None, [vp, PP, PP]),
        "orc_demod_user_filter": (None, [vp, PP, PP]),
        "orc_demod_fm": (None, [vp, PP, PP, C.c_float]), "orc_demod_am": (None, [vp, PP, PP]),
        "orc_demod_ssb": (None, [vp, PP, PP, C.c_int]), "orc_demod_cw": (None, [vp, PP, PP]),
        "orc_audiosink_new": (vp, [C.c_int, C.c_int]), "orc_audiosink_free": (None, [vp]),
        "orc_audiosink_filter": (C.c_int, [vp, PP, PP]),
        "orc_spectrum_run": (C.c_longlong, [C.c_int, _u8p, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
        "orc_chain_run": (C.c_longlong, [C.c_int, _u8p, C.c_longlong, C.c_int, C.c_longlong, C.c_longlong,
                                         C.c_int, C.c_int, C.c_int, C.c_float, _f32p, C.c_longlong]),
        "orc_synth_iq": (None, [C.c_int, C.c_uint32, C.POINTER(SynthComp), C.c_int, C.c_int, C.c_longlong,
                                C.c_longlong, _u8p]),
        "orc_synth_default_comps": (C.c_int, [C.c_int, C.POINTER(SynthComp)]),
        "orc_synth_step": (C.c_uint32, [C.c_double]),
        "orc_iqconv_new": (vp, [C.c_void_p, C.c_int]), "orc_iqconv_free": (None, [vp]), "orc_iqconv_reset": (None, [vp]),
        "orc_iqconv_process": (None, [vp, C.c_void_p, C.c_longlong]),
        "orc_airspy_convert_samples": (None, [C.c_void_p, C.c_void_p, C.c_longlong]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


_LIB = None
_REF = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = _declare(C.CDLL(so))
    return _LIB


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "librfa_ref.so"))


def ref():
    """The compiled reference (pffft.c + nativedsp.cpp) plus the restatement, one .so."""
    global _REF
    if _REF is None:
        so = os.path.join(HERE, "_ref", "librfa_ref.so")
        if not os.path.exists(so):
            build()
        L = _declare(C.CDLL(so))
        L.ref_perform_fft.restype, L.ref_perform_fft.argtypes = None, [_f32p, _f32p, C.c_int]
        L.ref_perform_fft_logmag.restype, L.ref_perform_fft_logmag.argtypes = None, [_f32p, _f32p, C.c_int]
        L.ref_pffft_simd_size.restype, L.ref_pffft_simd_size.argtypes = C.c_int, []
        L.ref_iqconv_new.restype, L.ref_iqconv_new.argtypes = C.c_void_p, [C.c_void_p, C.c_int]
        L.ref_iqconv_free.restype, L.ref_iqconv_free.argtypes = None, [C.c_void_p]
        L.ref_iqconv_process.restype, L.ref_iqconv_process.argtypes = None, [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_airspy_hb_kernel.restype, L.ref_airspy_hb_kernel.argtypes = C.c_int, [C.c_void_p, C.c_int]
        L.ref_spectrum_run.restype = C.c_longlong
        L.ref_spectrum_run.argtypes = [C.c_int, _u8p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]
        _REF = L
    return _REF


# ---------------------------------------------------------------- helpers ----
def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PacketView:
    """Owns an orc_packet and exposes numpy views of its re/im arrays."""

    def __init__(self, capacity, L=None):
        self.L = L or lib()
        self.p = self.L.orc_packet_new(int(capacity))
        cap = max(int(capacity), 1)
        self.re = np.ctypeslib.as_array(self.p.contents.re, shape=(cap,))
        self.im = np.ctypeslib.as_array(self.p.contents.im, shape=(cap,))

    def __del__(self):
        try:
            self.L.orc_packet_free(self.p)
        except Exception:
            pass

    @property
    def capacity(self):
        return self.p.contents.capacity

    @property
    def size(self):
        return self.p.contents.size

    @size.setter
    def size(self, v):
        self.p.contents.size = min(int(v), self.p.contents.capacity)

    @property
    def sampleRate(self):
        return self.p.contents.sampleRate

    @sampleRate.setter
    def sampleRate(self, v):
        self.p.contents.sampleRate = int(v)

    @property
    def frequency(self):
        return self.p.contents.frequency

    @frequency.setter
    def frequency(self, v):
        self.p.contents.frequency = int(v)

    def load(self, re, im=None, sampleRate=None):
        n = len(re)
        self.re[:n] = re
        if im is not None:
            self.im[:n] = im
        self.size = n
        if sampleRate is not None:
            self.sampleRate = sampleRate
        return self

    def out_re(self):
        return self.re[: self.size].copy()

    def out_im(self):
        return self.im[: self.size].copy()


def default_comps(fmt):
    arr = (SynthComp * 3)()
    n = lib().orc_synth_default_comps(fmt, arr)
    return [(arr[i].step, arr[i].amp, arr[i].modStep, arr[i].modK) for i in range(n)]


def synth_step(cycles_per_sample):
    return int(lib().orc_synth_step(float(cycles_per_sample)))


def synth_iq(fmt, nsamples, first=0, comps=None, noise_shift=2, seed=SEED):
    comps = default_comps(fmt) if comps is None else comps
    arr = (SynthComp * max(len(comps), 1))()
    for i, c in enumerate(comps):
        arr[i] = SynthComp(*[int(x) for x in c])
    out = np.empty(nsamples * BYTES_PER_SAMPLE[fmt], dtype=np.uint8)
    lib().orc_synth_iq(fmt, seed, arr, len(comps), noise_shift, first, nsamples, out)
    return out


def lowpass_taps(gain, fs, cutoff, tw, att, window=WIN_BLACKMAN, beta=0.0, max_taps=0):
    ptr = C.POINTER(C.c_float)()
    n = lib().orc_lowpass_taps(gain, fs, cutoff, tw, att, window, beta, max_taps, C.byref(ptr))
    if n == 0:
        return None
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def bandpass_taps(gain, fs, lo, hi, tw, att):
    pr, pi = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
    n = lib().orc_bandpass_taps(gain, fs, lo, hi, tw, att, C.byref(pr), C.byref(pi))
    if n == 0:
        return None, None
    return np.ctypeslib.as_array(pr, shape=(n,)).copy(), np.ctypeslib.as_array(pi, shape=(n,)).copy()


def resampler_taps(interp, decim, fractional_bw=0.4, max_taps=0):
    ptr = C.POINTER(C.c_float)()
    n = lib().orc_design_resampler_taps(interp, decim, fractional_bw, max_taps, C.byref(ptr))
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, np.float32)


def limit_denominator(num, den, max_den=10000):
    a, b = C.c_int(), C.c_int()
    lib().orc_limit_denominator(num, den, max_den, C.byref(a), C.byref(b))
    return a.value, b.value


def nativedsp_window(N):
    w = np.empty(N, dtype=np.float32)
    lib().orc_nativedsp_window(N, w)
    return w


def spectrum_run(fmt, iq, N, L=0):
    """Restated spectrum path over a whole recording -> rows[F][N], peaks[N], avg[N]."""
    ns = len(iq) // BYTES_PER_SAMPLE[fmt]
    F = ns // N
    rows = np.empty((F, N), dtype=np.float32)
    peaks = np.empty(N, dtype=np.float32)
    avg = np.empty(N, dtype=np.float32)
    lib().orc_spectrum_run(fmt, iq, ns, N, L, _vp(rows), _vp(peaks), _vp(avg))
    return rows, peaks, avg


class IqConverterInt16:
    """iqconverter_int16 (Airspy real -> IQ): the C restatement (`use_ref=False`) or the reference's own
    iqconverter_int16.c compiled in place (`use_ref=True`, oracle/_ref).  process() works in place like the
    reference and returns the array."""

    def __init__(self, hb_kernel, use_ref=False):
        self.kernel = np.ascontiguousarray(hb_kernel, np.int16)
        self.use_ref = use_ref
        self.L = ref() if use_ref else lib()
        new = self.L.ref_iqconv_new if use_ref else self.L.orc_iqconv_new
        self.h = new(self.kernel.ctypes.data, len(self.kernel))

    def process(self, samples):
        assert samples.dtype == np.int16 and samples.flags.c_contiguous and len(samples) % 4 == 0
        if self.use_ref:
            self.L.ref_iqconv_process(self.h, samples.ctypes.data, len(samples))
        else:
            self.L.orc_iqconv_process(self.h, samples.ctypes.data, len(samples))
        return samples

    def __del__(self):
        try:
            (self.L.ref_iqconv_free if self.use_ref else self.L.orc_iqconv_free)(self.h)
        except Exception:
            pass


def airspy_hb_kernel():
    """HB_KERNEL_INT16 of the reference's libairspy/filters.h, read out of oracle/_ref (the table is compiled into
    that library from the reference tree; it is not copied into this repository).  None when _ref is absent."""
    if not ref_available():
        return None
    n = ref().ref_airspy_hb_kernel(None, 0)
    k = np.zeros(n, np.int16)
    ref().ref_airspy_hb_kernel(k.ctypes.data, n)
    return k


def synthetic_hb_kernel(ntaps=47):
    """A half-band kernel of our own for boxes without oracle/_ref (windowed sinc, odd taps zero except the centre)."""
    m = np.arange(ntaps) - ntaps // 2
    h = np.sinc(m / 2.0) * np.hamming(ntaps) * 0.5
    k = np.round(h * 32768).astype(np.int64)
    k[(m % 2 == 0) & (m != 0)] = 0
    return np.clip(k, -32768, 32767).astype(np.int16)


def ema_rows(rows, alpha, init=None):
    """Exponential average of rows[F][N] in time order (orc_ema_rows; not in the reference)."""
    rows = np.ascontiguousarray(rows, np.float32)
    avg = np.empty(rows.shape[1], np.float32)
    init = None if init is None else np.ascontiguousarray(init, np.float32)
    lib().orc_ema_rows(rows, rows.shape[0], rows.shape[1], float(alpha), _vp(init), avg)
    return avg


def ref_spectrum_run(fmt, iq, N, L=0, nthreads=1, want_rows=True):
    """Same, with the reference's own pffft (and, single-threaded, its JNI log-mag loop)."""
    ns = len(iq) // BYTES_PER_SAMPLE[fmt]
    F = ns // N
    rows = np.empty((F, N), dtype=np.float32) if want_rows else None
    peaks = np.empty(N, dtype=np.float32)
    avg = np.empty(N, dtype=np.float32)
    ref().ref_spectrum_run(fmt, iq, ns, N, L, _vp(rows), _vp(peaks), _vp(avg) if want_rows else None, nthreads)
    return rows, peaks, avg


def chain_run(fmt, iq, fs, src_freq, chan_freq, mode, channel_width, packet_samples, volume=1.0):
    ns = len(iq) // BYTES_PER_SAMPLE[fmt]
    cap = int(ns * 48000.0 / fs * 2.5) + 4096
    audio = np.empty(cap, dtype=np.float32)
    n = lib().orc_chain_run(fmt, iq, ns, fs, src_freq, chan_freq, mode, channel_width, packet_samples, volume,
                            audio, cap)
    return audio[:n].copy()
