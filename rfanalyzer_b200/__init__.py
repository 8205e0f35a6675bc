"""rfanalyzer_b200 -- B200 (sm_100a) implementation of RF Analyzer's IQ->spectrum and
IQ->audio hot path, behind the reference's own DSP interface.

The compute lives in librfa_b200.so (hand-written CUDA, C ABI in include/rfa_b200.h).
This package is the host-side mirror of the reference classes that sit on that path
(IQConverter, NativeDsp, FftProcessor, FirFilter, ComplexFirFilter, RationalResampler,
Demodulator, AudioSink) plus thin batch objects (Context, SpectrumPlan, ChainPlan).
"""
from . import _lib
from ._lib import (FMT_S8, FMT_U8, FMT_S16LE, WIN_BLACKMAN_REF, WIN_HANN, WIN_RECT, MEM_HOST, MEM_DEVICE, AVG_BOXCAR, AVG_EMA,
                   MODE_OFF, MODE_AM, MODE_NFM, MODE_WFM, MODE_LSB, MODE_USB, MODE_CW, SUM_FMA, SUM_EXACT,
                   BYTES_PER_SAMPLE, RfaError)
from .engine import Context, SpectrumPlan, synth_iq, synth_step, default_synth_components

from . import dsp
from . import detect
from .dsp import (SamplePacket, Signed8BitIQConverter, Unsigned8BitIQConverter, Signed16BitIQConverter, NativeDsp,
                  FftProcessor, FftProcessorData, FirFilter, ComplexFirFilter, RationalResampler, Demodulator, AudioSink,
                  ChainPlan, FileIQSource, parse_recording_name, recording_file_name, IqConverterInt16, Scheduler,
                  airspy_convert_samples)

__all__ = ["Context", "SpectrumPlan", "RfaError", "FMT_S8", "FMT_U8", "FMT_S16LE", "WIN_BLACKMAN_REF",
           "WIN_HANN", "WIN_RECT", "MEM_HOST", "MEM_DEVICE", "BYTES_PER_SAMPLE"]
