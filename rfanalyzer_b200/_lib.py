"""ctypes binding of librfa_b200.so (the C ABI declared in include/rfa_b200.h).

There is no fallback of any kind: if the shared library is missing, or a call fails
(for instance because no B200 is present), an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RFA_B200_LIB: load another build of the same library (librfa_b200_lab.so with the experimental kernels, or a
# -D tuning variant); never a fallback.  This is the only environment variable the package reads -- the library
# itself reads none (its knobs are rfa_ctx_set_option).
LIB_PATH = os.environ.get("RFA_B200_LIB") or os.path.join(HERE, "lib", "librfa_b200.so")
LAB_LIB_PATH = os.path.join(HERE, "lib", "librfa_b200_lab.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = range(5)
MEM_HOST, MEM_DEVICE = 0, 1
FMT_S8, FMT_U8, FMT_S16LE = 0, 1, 2
WIN_BLACKMAN_REF, WIN_HANN, WIN_RECT = 0, 1, 2
AVG_BOXCAR, AVG_EMA = 0, 1
TAPWIN_BLACKMAN, TAPWIN_HAMMING, TAPWIN_KAISER = 0, 1, 2
MODE_OFF, MODE_AM, MODE_NFM, MODE_WFM, MODE_LSB, MODE_USB, MODE_CW = range(7)
SUM_FMA, SUM_EXACT = 0, 1
BYTES_PER_SAMPLE = {FMT_S8: 2, FMT_U8: 2, FMT_S16LE: 4}


class RfaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"librfa_b200 error {code}: {message}")
        self.code = code


class SpectrumDesc(C.Structure):
    _fields_ = [("format", C.c_int), ("fft_size", C.c_int), ("window", C.c_int), ("avg_len", C.c_int),
                ("peak_hold", C.c_int), ("avg_mode", C.c_int), ("ema_alpha", C.c_float)]


class SpectrumOut(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("row0", C.c_longlong), ("row_step", C.c_longlong),
                ("ring_rows", C.c_longlong), ("row_stride", C.c_longlong), ("history_rows", C.c_longlong),
                ("peaks", C.c_void_p), ("peaks_accumulate", C.c_int), ("avg", C.c_void_p),
                ("avg_accumulate", C.c_int)]


class RecordingInfo(C.Structure):
    _fields_ = [("file_format", C.c_int), ("frequency", C.c_longlong), ("sample_rate", C.c_longlong),
                ("have_format", C.c_int), ("have_frequency", C.c_int), ("have_sample_rate", C.c_int)]


class RenderDesc(C.Structure):
    _fields_ = [("fft_size", C.c_int), ("frequency", C.c_longlong), ("sample_rate", C.c_int),
                ("viewport_frequency", C.c_longlong), ("viewport_sample_rate", C.c_longlong),
                ("width", C.c_int), ("fft_height", C.c_int), ("min_db", C.c_float), ("max_db", C.c_float),
                ("avg_len", C.c_int), ("ring_rows", C.c_int), ("row_stride", C.c_longlong),
                ("newest_row", C.c_int), ("first_row", C.c_int), ("nrows", C.c_int)]


class SynthComp(C.Structure):
    _fields_ = [("step", C.c_uint32), ("amp", C.c_int32), ("mod_step", C.c_uint32), ("mod_k", C.c_int32)]


class ChainDesc(C.Structure):
    _fields_ = [("format", C.c_int), ("sample_rate", C.c_int), ("source_frequency", C.c_longlong),
                ("channel_frequency", C.c_longlong), ("mode", C.c_int), ("channel_width", C.c_int),
                ("packet_samples", C.c_int), ("volume", C.c_float), ("flags", C.c_int)]


class SchedulerDesc(C.Structure):
    _fields_ = [("format", C.c_int), ("sample_rate", C.c_int), ("source_frequency", C.c_longlong),
                ("packet_samples", C.c_int), ("fft_size", C.c_int), ("window", C.c_int), ("avg_len", C.c_int),
                ("peak_hold", C.c_int), ("ring_rows", C.c_int), ("demodulation_mode", C.c_int),
                ("channel_frequency", C.c_longlong), ("channel_width", C.c_int), ("volume", C.c_float), ("flags", C.c_int),
                ("squelch_enabled", C.c_int), ("squelch_db", C.c_float), ("record_only_when_squelch_satisfied", C.c_int)]


class SchedulerIO(C.Structure):
    _fields_ = [("packets", C.c_longlong), ("frames", C.c_longlong), ("signal_strength", C.c_void_p),
                ("demod_gate", C.c_void_p), ("record_gate", C.c_void_p), ("audio", C.c_void_p),
                ("audio_capacity", C.c_longlong), ("n_audio", C.c_longlong)]


class DetectWindow(C.Structure):
    """rfa_detect_window: (row, first bin, last bin inclusive)."""
    _fields_ = [("row", C.c_longlong), ("start", C.c_int), ("end", C.c_int)]


class Signal(C.Structure):
    """rfa_signal: DiscoveredSignal's numeric fields."""
    _fields_ = [("frequency", C.c_longlong), ("peak", C.c_float), ("average", C.c_float), ("bandwidth", C.c_longlong),
                ("grouped", C.c_int)]


DETECT_PEAK_ONLY, DETECT_AVERAGE_ONLY, DETECT_PEAK_OR_AVERAGE = 0, 1, 2

_vp, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_pi, _pll, _pvp = C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every function include/rfa_b200.h declares is listed here and
# tests/test_abi.py checks the two stay in step.
SIGNATURES = {
    "rfa_version": (_i, []),
    "rfa_last_error": (C.c_char_p, []),
    "rfa_ctx_create": (_i, [_i, _vp, _pvp]),
    "rfa_ctx_destroy": (_i, [_vp]),
    "rfa_ctx_sync": (_i, [_vp]),
    "rfa_ctx_device": (_i, [_vp]),
    "rfa_ctx_sm_count": (_i, [_vp]),
    "rfa_ctx_stream": (_vp, [_vp]),
    "rfa_ctx_launch_count": (_ll, [_vp]),
    "rfa_ctx_set_option": (_i, [_vp, C.c_char_p, _ll]),
    "rfa_ctx_get_option": (_i, [_vp, C.c_char_p, _pll]),
    "rfa_host_alloc": (_i, [C.c_size_t, _pvp]),
    "rfa_host_free": (_i, [_vp]),
    "rfa_convert": (_i, [_vp, _i, _vp, _ll, _vp, _vp, _i]),
    "rfa_nco_design": (_i, [_i, _i, _i, _pi, _pi, _vp, _vp]),
    "rfa_mix": (_i, [_vp, _i, _vp, _ll, _vp, _vp, _i, _i, _vp, _vp, _i]),
    "rfa_make_window": (_i, [_i, _i, _vp]),
    "rfa_fft_c2c": (_i, [_vp, _vp, _vp, _i, _ll, _i]),
    "rfa_fft_logmag": (_i, [_vp, _vp, _vp, _i, _ll, _i]),
    "rfa_windowed_fft_logmag": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _i, _i]),
    "rfa_spectrum_plan_create": (_i, [_vp, C.POINTER(SpectrumDesc), _pvp]),
    "rfa_spectrum_plan_destroy": (_i, [_vp]),
    "rfa_spectrum_process": (_i, [_vp, _vp, _ll, C.POINTER(SpectrumOut), _i]),
    "rfa_spectrum_algorithmic_bytes": (_ll, [_vp, _ll, _i]),
    "rfa_average_rows": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _ll, _i, _i, _vp, _i, _i]),
    "rfa_ema_rows": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _ll, _ll, _f, _i, _i, _vp, _i]),
    "rfa_channel_bins": (_i, [_i, _ll, _i, _ll, _ll, _pi, _pi]),
    "rfa_channel_strength": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _ll, _i, _i, _vp, _i]),
    "rfa_shift_rows": (_i, [_vp, _vp, _ll, _ll, _i, _i]),
    "rfa_detect_windows": (_i, [_vp, _vp, _ll, _i, _vp, _i, _vp, _vp, _i, _i]),
    "rfa_detect_bin": (_i, [_ll, _ll, _i, _ll]),
    "rfa_detect_half_width": (_i, [_ll, _i, _i, _i]),
    "rfa_detect_window_at": (_i, [_ll, _ll, _i, _ll, _i, _i, _pi, _pi, _pi]),
    "rfa_detect_decide": (_i, [_f, _f, _f, _f, _f, _i]),
    "rfa_scan_grid": (_ll, [_ll, _ll, _ll, _ll, _ll, _ll, _i, _ll, _vp, _vp, _ll]),
    "rfa_group_signals": (_ll, [_vp, _ll, _ll, _i, _vp]),
    "rfa_fill": (_i, [_vp, _vp, _ll, _f]),
    "rfa_tap_window": (_i, [_i, _d, _i, _i, C.POINTER(C.c_float)]),
    "rfa_design_lowpass": (_i, [_f, _f, _f, _f, _f, _i, _d, _i, _vp, _i, _pi]),
    "rfa_design_bandpass": (_i, [_f, _f, _f, _f, _f, _f, _vp, _vp, _i, _pi]),
    "rfa_limit_denominator": (_i, [_i, _i, _i, _pi, _pi]),
    "rfa_design_resampler_taps": (_i, [_i, _i, _f, _i, _vp, _i, _pi]),
    "rfa_fir_create": (_i, [_vp, _vp, _vp, _i, _i, _i, _pvp]),
    "rfa_fir_destroy": (_i, [_vp]),
    "rfa_fir_reset": (_i, [_vp]),
    "rfa_fir_process": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _ll, _pll, _pll, _i]),
    "rfa_resampler_create": (_i, [_vp, _i, _i, _vp, _i, _f, _i, _i, _pvp]),
    "rfa_resampler_destroy": (_i, [_vp]),
    "rfa_resampler_info": (_i, [_vp, _pi, _pi, _pi]),
    "rfa_resampler_process": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _ll, _pll, _pll, _i]),
    "rfa_demod_fm": (_i, [_vp, _vp, _vp, _ll, _vp, _f, _f, _vp, _i, _i]),
    "rfa_demod_am": (_i, [_vp, _vp, _vp, _ll, _vp, _f, _vp, _i, _i]),
    "rfa_agc": (_i, [_vp, _vp, _ll, _vp, _f, _i, _i]),
    "rfa_mode_info": (_i, [_i, _pi, _pi, _pi, _pi]),
    "rfa_spectrum_plan_info": (_i, [_vp, _pi, _pi, _pi]),
    "rfa_recording_parse_name": (_i, [C.c_char_p, C.POINTER(RecordingInfo)]),
    "rfa_recording_file_name": (_i, [C.c_char_p, C.c_char_p, _i, _ll, _ll, C.c_char_p, _i]),
    "rfa_recording_sample_format": (_i, [_i]),
    "rfa_file_source_open": (_i, [C.c_char_p, _i, _ll, _i, _ll, _pvp]),
    "rfa_file_source_get_packet": (_i, [_vp, _vp]),
    "rfa_file_source_bytes_read": (_ll, [_vp]),
    "rfa_file_source_close": (_i, [_vp]),
    "rfa_spectrum_process_file": (_i, [_vp, C.c_char_p, _ll, _ll, C.POINTER(SpectrumOut), _ll, _pll]),
    "rfa_render_waterfall": (_i, [_vp, C.POINTER(RenderDesc), _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "rfa_chain_create": (_i, [_vp, C.POINTER(ChainDesc), _pvp]),
    "rfa_chain_destroy": (_i, [_vp]),
    "rfa_chain_info": (_i, [_vp, _pi, _pi, _pi, _pi, _pi, _pi, _pi]),
    "rfa_chain_max_audio": (_ll, [_vp, _ll]),
    "rfa_chain_process": (_i, [_vp, _vp, _ll, _vp, _ll, _pll, _i]),
    "rfa_chain_seek": (_i, [_vp, _ll, _pll]),
    "rfa_scheduler_create": (_i, [_vp, C.POINTER(SchedulerDesc), _pvp]),
    "rfa_scheduler_destroy": (_i, [_vp]),
    "rfa_scheduler_process": (_i, [_vp, _vp, _ll, C.POINTER(SchedulerIO), _i]),
    "rfa_scheduler_state": (_i, [_vp, _pvp, _pll, _pll, _pvp, _pvp, _pi, _pi, _pll, _pll]),
    "rfa_scheduler_read": (_i, [_vp, _vp, _vp, _vp]),
    "rfa_iqconverter_create": (_i, [_vp, _vp, _i, _pvp]),
    "rfa_iqconverter_destroy": (_i, [_vp]),
    "rfa_iqconverter_reset": (_i, [_vp]),
    "rfa_iqconverter_process": (_i, [_vp, _vp, _ll, _i]),
    "rfa_iqconverter_stats": (_i, [_vp, _pll, _pll, _pll]),
    "rfa_airspy_convert_samples": (_i, [_vp, _vp, _vp, _ll, _i]),
    "rfa_synth_iq": (_i, [_vp, _i, C.c_uint32, C.POINTER(SynthComp), _i, _i, _ll, _ll, _vp, _i]),
}

_LIB = None


def load():
    """Load librfa_b200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RfaError(-1, f"{LIB_PATH} is missing: build it with `make -C rfanalyzer_b200/csrc` "
                               "(there is no fallback implementation)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


def check(rc):
    if rc != OK:
        raise RfaError(rc, load().rfa_last_error().decode("utf-8", "replace"))


def ptr(x):
    """Address of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x.ctypes.data
