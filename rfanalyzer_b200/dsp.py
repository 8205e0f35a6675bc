"""Host-side mirror of the reference's DSP classes over the C ABI.

Same names, argument meaning and return values as the reference (so tests read like its
own androidTest sources), but every sample is processed on the GPU by librfa_b200:

  SamplePacket                A/source/SamplePacket.java
  Signed8BitIQConverter ...   A/source/IQConverter.java and its three subclasses
  NativeDsp                   nativedsp/src/main/java/com/mantz_it/nativedsp/NativeDsp.kt
  FftProcessor(+Data)         A/analyzer/FftProcessor.kt (one loop iteration per process())
  FirFilter, ComplexFirFilter A/dsp/FirFilter.kt, A/dsp/ComplexFirFilter.java
  RationalResampler           A/dsp/RationalResampler.kt
  Demodulator, AudioSink      A/analyzer/Demodulator.kt, A/analyzer/AudioSink.java (DSP parts)
  ChainPlan                   the whole IQ -> audio chain over a recording, in one call

(A/ = app/src/main/java/com/mantz_it/rfanalyzer/.)  Packets live in host memory like the JVM's
float[]; the batch objects (SpectrumPlan, ChainPlan) are the fast path.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr
from .engine import SpectrumPlan

AUDIO_RATE = 48000


class SamplePacket:
    """SamplePacket.java:28-137: planar float re/im with a fill level."""

    def __init__(self, size_or_re, im=None, frequency=0, sampleRate=0, size=None):
        if im is None:
            n = int(size_or_re)
            self._re = np.zeros(n, np.float32)
            self._im = np.zeros(n, np.float32)
            self._size = 0
        else:
            self._re = np.ascontiguousarray(size_or_re, np.float32)
            self._im = np.ascontiguousarray(im, np.float32)
            if len(self._re) != len(self._im):
                raise ValueError("Arrays must be of the same length")
            self._size = len(self._re) if size is None else int(size)
            if self._size > len(self._re):
                raise ValueError("Size must be of the smaller or equal the array length")
        self.frequency = int(frequency)
        self.sampleRate = int(sampleRate)

    def re(self, i=None):
        return self._re if i is None else self._re[i]

    def im(self, i=None):
        return self._im if i is None else self._im[i]

    def capacity(self):
        return len(self._re)

    def size(self):
        return self._size

    def setSize(self, size):
        self._size = min(int(size), len(self._re))


# ------------------------------------------------------------------------------ converters
class IQConverter:
    """IQConverter.java:31-87.  Subclasses set FMT."""
    FMT = None
    MAX_COSINE_LENGTH = 500

    def __init__(self, ctx):
        self.ctx = ctx
        self.frequency = 0
        self.sampleRate = 0
        self.cosineFrequency = 0
        self.cosineIndex = 0
        self._cos = self._sin = None

    def setFrequency(self, frequency):
        self.frequency = int(frequency)

    def setSampleRate(self, sampleRate):
        if self.sampleRate != sampleRate:
            self.sampleRate = int(sampleRate)
            self.cosineFrequency = -1

    def _bytes_per_sample(self):
        return _lib.BYTES_PER_SAMPLE[self.FMT]

    def _room(self, packet, samplePacket):
        start = samplePacket.size()
        if start >= samplePacket.capacity():
            return start, 0
        avail = len(packet) // self._bytes_per_sample()
        return start, min(avail, samplePacket.capacity() - start)

    def fillPacketIntoSamplePacket(self, packet, samplePacket):
        packet = np.ascontiguousarray(packet, np.uint8)
        start, count = self._room(packet, samplePacket)
        if count == 0:
            return 0
        re = samplePacket.re()[start:start + count]
        im = samplePacket.im()[start:start + count]
        self.ctx.convert(self.FMT, packet, count, re, im)
        samplePacket.setSize(start + count)
        samplePacket.sampleRate = self.sampleRate
        samplePacket.frequency = self.frequency
        return count

    def generateMixerLookupTable(self, mixFrequency):
        amix = abs(mixFrequency)
        if mixFrequency == 0 or self.sampleRate // amix > self.MAX_COSINE_LENGTH:
            mixFrequency += self.sampleRate
        if self._cos is None or mixFrequency != self.cosineFrequency:
            eff, self._cos, self._sin = self.ctx.nco_design(self.FMT, self.sampleRate, mixFrequency)
            self.cosineFrequency = mixFrequency
            self.cosineIndex = 0

    def mixPacketIntoSamplePacket(self, packet, samplePacket, channelFrequency):
        mix = ((int(self.frequency) - int(channelFrequency) + 2 ** 31) % 2 ** 32) - 2 ** 31  # Java (int) of a long
        self.generateMixerLookupTable(mix)
        packet = np.ascontiguousarray(packet, np.uint8)
        start, count = self._room(packet, samplePacket)
        if count == 0 or len(self._cos) == 0:
            return 0
        if self.cosineIndex >= len(self._cos):
            self.cosineIndex = 0
        re = samplePacket.re()[start:start + count]
        im = samplePacket.im()[start:start + count]
        self.ctx.mix(self.FMT, packet, count, self._cos, self._sin, self.cosineIndex, re, im)
        self.cosineIndex = (self.cosineIndex + count) % len(self._cos)
        samplePacket.setSize(start + count)
        samplePacket.sampleRate = self.sampleRate
        samplePacket.frequency = int(channelFrequency)
        return count


class Signed8BitIQConverter(IQConverter):
    FMT = _lib.FMT_S8


class Unsigned8BitIQConverter(IQConverter):
    FMT = _lib.FMT_U8


class Signed16BitIQConverter(IQConverter):
    FMT = _lib.FMT_S16LE


# ------------------------------------------------------------------------------ NativeDsp
class NativeDsp:
    """NativeDsp.kt: Blackman window + FFT + log magnitude, fft-shifted."""

    def __init__(self, ctx):
        self.ctx = ctx

    def performWindowedFftAndReturnMag(self, re, im, magOut):
        n = len(re)
        if len(im) != n or len(magOut) != n:
            return False
        self.ctx.windowed_fft_logmag(np.ascontiguousarray(re, np.float32), np.ascontiguousarray(im, np.float32),
                                     magOut, n, 1)
        return True


class FftProcessorData:
    """FftProcessor.kt:43-61; the waterfall ring and the peaks live in GPU memory."""

    def __init__(self):
        self.waterfallBuffer = None  # torch [rows, N] on the device
        self.peaks = None
        self.frequencyOrSampleRateChanged = True
        self.writeIndex = 0
        self.readIndex = 0
        self.frequency = None
        self.sampleRate = None
        self.rowsValid = 0


class FftProcessor:
    """One iteration of FftProcessor.run (FftProcessor.kt:111-253) per process(): FFT, channel signal
    strength, ring maintenance (incl. the history shift on retune), peak hold."""
    BUFFER_SIZES = {"SLOW": 500, "NORMAL": 400, "FAST": 300}

    def __init__(self, ctx, fftProcessorData, waterfallSpeed="NORMAL", fftPeakHold=False, fmt=None,
                 getChannelFrequencyRange=None, onAverageSignalStrengthChanged=None, avg_len=0):
        import torch
        self.torch = torch
        self.ctx, self.data = ctx, fftProcessorData
        self.waterfallSpeed, self.fftPeakHold = waterfallSpeed, fftPeakHold
        self.getChannelFrequencyRange = getChannelFrequencyRange or (lambda: None)
        self.onAverageSignalStrengthChanged = onAverageSignalStrengthChanged or (lambda v: None)
        self.avg_len = avg_len
        self.lastFrequency = self.lastSampleRate = None
        self._plans = {}

    def _plan(self, fmt, n):
        key = (fmt, n, self.avg_len)
        if key not in self._plans:
            self._plans[key] = SpectrumPlan(self.ctx, fmt, n, avg_len=self.avg_len, peak_hold=True)
        return self._plans[key]

    def process_iq(self, fmt, iq, nframes, fft_size, frequency, sampleRate, avg=None):
        """Batched equivalent of `nframes` loop iterations fed with consecutive raw-IQ FFT frames."""
        torch, d = self.torch, self.data
        n = fft_size
        ring = self.BUFFER_SIZES[self.waterfallSpeed]
        frequencyChanged = frequency != self.lastFrequency
        sampleRateChanged = sampleRate != self.lastSampleRate
        d.frequencyOrSampleRateChanged = frequencyChanged or sampleRateChanged
        frequencyDiff = (self.lastFrequency - frequency) if self.lastFrequency is not None else 0
        samplesPerHz = np.float32(n) / np.float32(sampleRate)
        self.lastFrequency, self.lastSampleRate = frequency, sampleRate
        d.frequency, d.sampleRate = frequency, sampleRate
        if d.waterfallBuffer is None or d.waterfallBuffer.shape[1] != n or d.waterfallBuffer.shape[0] != ring:
            d.waterfallBuffer = torch.full((ring, n), -9999.0, dtype=torch.float32, device="cuda")
            d.writeIndex, d.rowsValid = 0, ring  # every row holds -9999f, i.e. is "valid" history
        if frequencyDiff != 0:  # FftProcessor.kt:199-217
            shift = int(np.float32(frequencyDiff) * samplesPerHz)
            if abs(shift) < n:
                self.ctx.shift_rows(d.waterfallBuffer, ring, n, n, shift)
            else:
                d.waterfallBuffer.fill_(-9999.0)
        elif sampleRateChanged:
            d.waterfallBuffer.fill_(-9999.0)
        accumulate = True
        if self.fftPeakHold:
            if d.peaks is None or d.peaks.shape[0] != n or d.frequencyOrSampleRateChanged:
                d.peaks = torch.full((n,), -999999.0, dtype=torch.float32, device="cuda")
        else:
            d.peaks = None
        self._plan(fmt, n).process(iq, nframes, rows=d.waterfallBuffer, peaks=d.peaks, avg=avg, row0=d.writeIndex,
                                   row_step=-1, ring_rows=ring, history_rows=ring, peaks_accumulate=accumulate)
        d.readIndex = (d.writeIndex - (nframes - 1)) % ring
        d.writeIndex = (d.writeIndex - nframes) % ring
        rng = self.getChannelFrequencyRange()
        if rng is not None:  # FftProcessor.kt:143-157, for the newest row
            b0, b1 = self.ctx.channel_bins(n, frequency, sampleRate, rng[0], rng[1])
            if b1 > b0:
                out = np.zeros(1, np.float32)
                self.ctx.channel_strength(d.waterfallBuffer, d.readIndex, 1, ring, n, 1, b0, b1, out)
                self.onAverageSignalStrengthChanged(float(out[0]))


# ------------------------------------------------------------------------------ filters
def _taps_call(fn, *args):
    n = C.c_int()
    buf = np.zeros(1 << 16, np.float32)
    check(fn(*args, ptr(buf), len(buf), C.byref(n)))
    return buf[: n.value].copy()


class FirFilter:
    """FirFilter.kt:34-263."""

    def __init__(self, ctx, taps, decimation, gain=1.0, sampleRate=1.0, cutOffFrequency=0.0, transitionWidth=0.0,
                 attenuation=0.0, flags=_lib.SUM_EXACT):
        self.ctx = ctx
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.decimation, self.gain, self.sampleRate = decimation, gain, sampleRate
        self.cutOffFrequency, self.transitionWidth, self.attenuation = cutOffFrequency, transitionWidth, attenuation
        self.handle = C.c_void_p()
        check(ctx.lib.rfa_fir_create(ctx.handle, ptr(self.taps), None, len(self.taps), int(decimation), flags,
                                     C.byref(self.handle)))

    def __del__(self):
        try:
            self.ctx.lib.rfa_fir_destroy(self.handle)
        except Exception:
            pass

    @property
    def numberOfTaps(self):
        return len(self.taps)

    @staticmethod
    def createLowPassTaps(decimation, gain, sampleRate, cutoffFrequency, transitionWidth, attenuationInDecibels,
                          window=_lib.TAPWIN_BLACKMAN, beta=0.0, maxTaps=0):
        lib = _lib.load()
        n = C.c_int()
        buf = np.zeros(1 << 20, np.float32)
        rc = lib.rfa_design_lowpass(gain, sampleRate, cutoffFrequency, transitionWidth, attenuationInDecibels, window,
                                    beta, maxTaps, ptr(buf), len(buf), C.byref(n))
        return None if rc != 0 else buf[: n.value].copy()

    @classmethod
    def createLowPass(cls, ctx, decimation, gain, sampleRate, cutoffFrequency, transitionWidth,
                      attenuationInDecibels, flags=_lib.SUM_EXACT):
        taps = cls.createLowPassTaps(decimation, gain, sampleRate, cutoffFrequency, transitionWidth,
                                     attenuationInDecibels)
        if taps is None:
            return None
        return cls(ctx, taps, decimation, gain, sampleRate, cutoffFrequency, transitionWidth, attenuationInDecibels, flags)

    def _run(self, inPacket, outPacket, offset, length, real):
        start = outPacket.size()
        room = outPacket.capacity() - start
        nout, cons = C.c_longlong(), C.c_longlong()
        in_re = inPacket.re()[offset:offset + length]
        in_im = None if real else inPacket.im()[offset:offset + length]
        out_re = outPacket.re()[start:]
        out_im = None if real else outPacket.im()[start:]
        check(self.ctx.lib.rfa_fir_process(self.handle, ptr(in_re), ptr(in_im), length, ptr(out_re), ptr(out_im),
                                           room, C.byref(nout), C.byref(cons), _lib.MEM_HOST))
        outPacket.setSize(start + nout.value)
        outPacket.sampleRate = inPacket.sampleRate // self.decimation
        return cons.value

    def filter(self, inPacket, outPacket, offset, length):
        return self._run(inPacket, outPacket, offset, length, False)

    def filterReal(self, inPacket, outPacket, offset, length):
        return self._run(inPacket, outPacket, offset, length, True)


class ComplexFirFilter:
    """ComplexFirFilter.java:33-279."""

    def __init__(self, ctx, tapsReal, tapsImag, decimation, gain, sampleRate, lowCutOffFrequency,
                 highCutOffFrequency, transitionWidth, attenuation, flags=_lib.SUM_EXACT):
        self.ctx = ctx
        self.tapsReal = np.ascontiguousarray(tapsReal, np.float32)
        self.tapsImag = np.ascontiguousarray(tapsImag, np.float32)
        self.decimation, self.gain, self.sampleRate = decimation, gain, sampleRate
        self.lowCutOffFrequency, self.highCutOffFrequency = lowCutOffFrequency, highCutOffFrequency
        self.transitionWidth, self.attenuation = transitionWidth, attenuation
        self.handle = C.c_void_p()
        check(ctx.lib.rfa_fir_create(ctx.handle, ptr(self.tapsReal), ptr(self.tapsImag), len(self.tapsReal),
                                     int(decimation), flags, C.byref(self.handle)))

    def __del__(self):
        try:
            self.ctx.lib.rfa_fir_destroy(self.handle)
        except Exception:
            pass

    def getNumberOfTaps(self):
        return len(self.tapsReal)

    @classmethod
    def createBandPass(cls, ctx, decimation, gain, sampling_freq, low_cutoff_freq, high_cutoff_freq, transition_width,
                       attenuation_dB, flags=_lib.SUM_EXACT):
        lib = _lib.load()
        n = C.c_int()
        tre, tim = np.zeros(1 << 16, np.float32), np.zeros(1 << 16, np.float32)
        rc = lib.rfa_design_bandpass(gain, sampling_freq, low_cutoff_freq, high_cutoff_freq, transition_width,
                                     attenuation_dB, ptr(tre), ptr(tim), len(tre), C.byref(n))
        if rc != 0:
            return None
        return cls(ctx, tre[: n.value], tim[: n.value], decimation, gain, sampling_freq, low_cutoff_freq,
                   high_cutoff_freq, transition_width, attenuation_dB, flags)

    def filter(self, inPacket, outPacket, offset, length):
        start = outPacket.size()
        nout, cons = C.c_longlong(), C.c_longlong()
        check(self.ctx.lib.rfa_fir_process(self.handle, ptr(inPacket.re()[offset:offset + length]),
                                           ptr(inPacket.im()[offset:offset + length]), length,
                                           ptr(outPacket.re()[start:]), ptr(outPacket.im()[start:]),
                                           outPacket.capacity() - start, C.byref(nout), C.byref(cons), _lib.MEM_HOST))
        outPacket.setSize(start + nout.value)
        outPacket.sampleRate = inPacket.sampleRate // self.decimation
        return cons.value


class RationalResampler:
    """RationalResampler.kt:36-257."""

    def __init__(self, ctx, interpolation, decimation, taps=None, fractionalBw=0.4, maxTaps=0, flags=_lib.SUM_EXACT):
        if interpolation <= 0:
            raise ValueError("Interpolation must be > 0")
        if decimation <= 0:
            raise ValueError("Decimation must be > 0")
        self.ctx = ctx
        self.handle = C.c_void_p()
        t = None if taps is None else np.ascontiguousarray(taps, np.float32)
        check(ctx.lib.rfa_resampler_create(ctx.handle, int(interpolation), int(decimation), ptr(t),
                                           0 if t is None else len(t), float(fractionalBw), int(maxTaps), flags,
                                           C.byref(self.handle)))
        i, d, nt = C.c_int(), C.c_int(), C.c_int()
        check(ctx.lib.rfa_resampler_info(self.handle, C.byref(i), C.byref(d), C.byref(nt)))
        self.interpolation, self.decimation, self.tapsPerPhase = i.value, d.value, nt.value

    def __del__(self):
        try:
            self.ctx.lib.rfa_resampler_destroy(self.handle)
        except Exception:
            pass

    def getInterpolation(self):
        return self.interpolation

    def getDecimation(self):
        return self.decimation

    @staticmethod
    def limitDenominator(numerator, denominator, maxDenominator=10000):
        a, b = C.c_int(), C.c_int()
        check(_lib.load().rfa_limit_denominator(int(numerator), int(denominator), int(maxDenominator), C.byref(a), C.byref(b)))
        return a.value, b.value

    @staticmethod
    def designResamplerTaps(interpolation, decimation, fractionalBw, maxTaps):
        lib = _lib.load()
        n = C.c_int()
        buf = np.zeros(1 << 22, np.float32)
        check(lib.rfa_design_resampler_taps(interpolation, decimation, fractionalBw, maxTaps, ptr(buf), len(buf), C.byref(n)))
        return buf[: n.value].copy()

    def resample(self, inPacket, outPacket, offset, length):
        start = outPacket.size()
        nout, cons = C.c_longlong(), C.c_longlong()
        check(self.ctx.lib.rfa_resampler_process(self.handle, ptr(inPacket.re()[offset:offset + length]),
                                                 ptr(inPacket.im()[offset:offset + length]), length,
                                                 ptr(outPacket.re()[start:]), ptr(outPacket.im()[start:]),
                                                 outPacket.capacity() - start, C.byref(nout), C.byref(cons),
                                                 _lib.MEM_HOST))
        outPacket.setSize(start + nout.value)
        outPacket.sampleRate = int(inPacket.sampleRate * self.interpolation // self.decimation)
        outPacket.frequency = inPacket.frequency
        return cons.value


# ------------------------------------------------------------------------------ demodulation
def mode_info(mode):
    q, lo, hi, de = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(_lib.load().rfa_mode_info(mode, C.byref(q), C.byref(lo), C.byref(hi), C.byref(de)))
    return q.value, lo.value, hi.value, de.value


class AudioSink:
    """The decimating audio filters of AudioSink.java:94-96 and applyAudioFilter (:215-237)."""

    def __init__(self, ctx, packetSize, sampleRate=AUDIO_RATE, flags=_lib.SUM_EXACT):
        self.sampleRate = sampleRate
        self.audioFilter1 = FirFilter.createLowPass(ctx, 2, 1.0, 1.0, 0.1, 0.15, 30.0, flags)
        self.audioFilter2 = FirFilter.createLowPass(ctx, 4, 1.0, 1.0, 0.1, 0.1, 30.0, flags)
        self.tmpAudioSamples = SamplePacket(packetSize)

    def applyAudioFilter(self, input, output):
        ratio = input.sampleRate // self.sampleRate
        if ratio == 8:
            self.tmpAudioSamples.setSize(0)
            self.audioFilter1.filterReal(input, self.tmpAudioSamples, 0, input.size())
            output.setSize(0)
            self.audioFilter2.filterReal(self.tmpAudioSamples, output, 0, self.tmpAudioSamples.size())
            return True
        if ratio == 2:
            output.setSize(0)
            self.audioFilter1.filterReal(input, output, 0, input.size())
            return True
        return False


class Demodulator:
    """Demodulator.kt: applyUserFilter + demodulate{FM,AM,SSB,CW} + volume, one packet per process()."""
    BAND_PASS_ATTENUATION = 40
    USER_FILTER_ATTENUATION = 60
    CW_OFFSET_FREQUENCY = 750

    def __init__(self, ctx, packetSize, flags=_lib.SUM_EXACT):
        self.ctx, self.flags = ctx, flags
        self.quadratureSamples = SamplePacket(packetSize)
        self.userFilter = None
        self.bandPassFilter = None
        self.carryOver = np.zeros(2, np.float32)
        self.lastMax = np.zeros(1, np.float32)
        self.audioVolumeLevel = 1.0
        self._mode = _lib.MODE_OFF
        self._channelWidth = 0

    @property
    def demodulationMode(self):
        return self._mode

    @demodulationMode.setter
    def demodulationMode(self, mode):
        self._mode = mode
        self.channelWidth = mode_info(mode)[3]

    @property
    def channelWidth(self):
        return self._channelWidth

    @channelWidth.setter
    def channelWidth(self, value):
        _, lo, hi, _ = mode_info(self._mode)
        self._channelWidth = max(lo, min(hi, int(value)))

    @property
    def quadratureRate(self):
        return mode_info(self._mode)[0]

    def applyUserFilter(self, input, output):
        if self.userFilter is None or int(self.userFilter.cutOffFrequency) != self.channelWidth:
            self.userFilter = FirFilter.createLowPass(self.ctx, 1, 1.0, float(input.sampleRate), float(self.channelWidth),
                                                      float(np.float32(input.sampleRate) * np.float32(0.10)),
                                                      float(self.USER_FILTER_ATTENUATION), self.flags)
            if self.userFilter is None:
                return
        output.setSize(0)
        self.userFilter.filter(input, output, 0, input.size())

    def process(self, inputSamples, audioBuffer):
        q = self.quadratureSamples
        self.applyUserFilter(inputSamples, q)
        audioBuffer.setSize(0)
        lib, h, m = self.ctx.lib, self.ctx.handle, self._mode
        n = q.size()
        vol = float(self.audioVolumeLevel)
        if m in (_lib.MODE_NFM, _lib.MODE_WFM):
            maxDev = np.float32(self.channelWidth) * np.float32(0.75 if m == _lib.MODE_NFM else 0.85)
            gain = np.float32(self.quadratureRate) / np.float32(2 * np.pi * float(maxDev))
            if n:
                check(lib.rfa_demod_fm(h, ptr(q.re()), ptr(q.im()), n, ptr(self.carryOver), float(gain), vol,
                                       ptr(audioBuffer.re()), self.flags, _lib.MEM_HOST))
            audioBuffer.setSize(n)
            audioBuffer.sampleRate = self.quadratureRate
        elif m == _lib.MODE_AM:
            if n:
                check(lib.rfa_demod_am(h, ptr(q.re()), ptr(q.im()), n, ptr(self.lastMax), vol, ptr(audioBuffer.re()),
                                       self.flags, _lib.MEM_HOST))
            audioBuffer.setSize(n)
            audioBuffer.sampleRate = self.quadratureRate
        elif m in (_lib.MODE_LSB, _lib.MODE_USB, _lib.MODE_CW):
            cw = self.channelWidth
            if m == _lib.MODE_CW:
                need = self.bandPassFilter is None or int(self.bandPassFilter.highCutOffFrequency) != self.CW_OFFSET_FREQUENCY + cw // 2
                if need:
                    self.bandPassFilter = ComplexFirFilter.createBandPass(
                        self.ctx, 1, 1.0, float(q.sampleRate), self.CW_OFFSET_FREQUENCY - cw / 2.0,
                        self.CW_OFFSET_FREQUENCY + cw / 2.0, float(np.float32(q.sampleRate) * np.float32(0.01)),
                        float(self.BAND_PASS_ATTENUATION), self.flags)
            else:
                upper = m == _lib.MODE_USB
                bp = self.bandPassFilter
                need = bp is None or (upper and int(bp.highCutOffFrequency) != cw) or (not upper and int(bp.lowCutOffFrequency) != -cw)
                if need:
                    self.bandPassFilter = ComplexFirFilter.createBandPass(
                        self.ctx, 2, 1.0, float(q.sampleRate), 200.0 if upper else -float(cw), float(cw) if upper else -200.0,
                        float(np.float32(q.sampleRate) * np.float32(0.01)), float(self.BAND_PASS_ATTENUATION), self.flags)
            if self.bandPassFilter is None:
                return
            self.bandPassFilter.filter(q, audioBuffer, 0, n)
            if audioBuffer.size():
                check(lib.rfa_agc(h, ptr(audioBuffer.re()), audioBuffer.size(), ptr(self.lastMax), vol, self.flags,
                                  _lib.MEM_HOST))


class ChainPlan:
    """The whole IQ -> 48 kHz audio chain over a recording (or a run of whole packets) in one call."""

    def __init__(self, ctx, fmt, sample_rate, source_frequency, channel_frequency, mode, channel_width=0,
                 packet_samples=65536, volume=1.0, flags=_lib.SUM_FMA):
        self.ctx = ctx
        self.desc = _lib.ChainDesc(fmt, int(sample_rate), int(source_frequency), int(channel_frequency), int(mode),
                                   int(channel_width), int(packet_samples), float(volume), int(flags))
        self.handle = C.c_void_p()
        check(ctx.lib.rfa_chain_create(ctx.handle, C.byref(self.desc), C.byref(self.handle)))
        v = [C.c_int() for _ in range(7)]
        check(ctx.lib.rfa_chain_info(self.handle, *[C.byref(x) for x in v]))
        (self.interpolation, self.decimation, self.taps_per_phase, self.quadrature_rate, self.channel_width,
         self.nco_length, self.nco_frequency) = [x.value for x in v]

    def close(self):
        if self.handle:
            self.ctx.lib.rfa_chain_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def max_audio(self, nsamples):
        return self.ctx.lib.rfa_chain_max_audio(self.handle, int(nsamples))

    def seek(self, sample_index):
        """Position the chain at a packet boundary of the recording (rfa_chain_seek); returns the number
        of audio samples a sequential run has produced before that point."""
        n = C.c_longlong()
        check(self.ctx.lib.rfa_chain_seek(self.handle, int(sample_index), C.byref(n)))
        return n.value

    def process(self, iq, nsamples, audio):
        """audio: float32 buffer of at least max_audio(nsamples); returns the number of samples written.
        Host arrays: the audio is there on return.  Device tensors: the call only enqueues work on the context's
        stream (the count is exact on return); synchronise the context or keep working on that stream."""
        n = C.c_longlong()
        from .engine import _mem_of
        cap = audio.numel() if hasattr(audio, "numel") else len(audio)
        check(self.ctx.lib.rfa_chain_process(self.handle, ptr(iq), int(nsamples), ptr(audio), int(cap), C.byref(n),
                                             _mem_of(iq, audio)))
        return n.value


class Scheduler:
    """Scheduler.run (A/analyzer/Scheduler.kt:140-298) over batches of packets: rfa_scheduler_*.  Field names follow
    the reference (channelFrequency, isDemodulationActivated, squelchSatisfied)."""
    SQUELCH_DEBOUNCE_COUNT = 50

    def __init__(self, ctx, fmt, sampleRate, frequency, packetSamples, fftSize, window=_lib.WIN_BLACKMAN_REF, avg_len=0,
                 peak_hold=True, ring_rows=300, mode=_lib.MODE_OFF, channelFrequency=0, channelWidth=0, volume=1.0,
                 flags=_lib.SUM_FMA, squelchEnabled=False, squelch=-30.0, recordOnlyWhenSquelchIsSatisfied=False):
        self.ctx = ctx
        self.desc = _lib.SchedulerDesc(int(fmt), int(sampleRate), int(frequency), int(packetSamples), int(fftSize), int(window),
                                       int(avg_len), 1 if peak_hold else 0, int(ring_rows), int(mode), int(channelFrequency),
                                       int(channelWidth), float(volume), int(flags), 1 if squelchEnabled else 0,
                                       float(squelch), 1 if recordOnlyWhenSquelchIsSatisfied else 0)
        self.handle = C.c_void_p()
        check(ctx.lib.rfa_scheduler_create(ctx.handle, C.byref(self.desc), C.byref(self.handle)))
        self.isDemodulationActivated = mode != _lib.MODE_OFF

    def max_audio(self, npackets):
        n = npackets * self.desc.packet_samples
        return int(n * 48000.0 / self.desc.sample_rate) + 64 * (npackets + 1)

    def process(self, packets, npackets, audio=None):
        """-> dict(frames, signal_strength[frames], demod_gate[npackets], record_gate[npackets], n_audio)."""
        from .engine import _mem_of
        max_frames = npackets * max(1, self.desc.packet_samples // self.desc.fft_size + 1) + 1
        strength = np.empty(max_frames, np.float32)
        dem, rec = np.zeros(npackets, np.uint8), np.zeros(npackets, np.uint8)
        cap = 0 if audio is None else (audio.numel() if hasattr(audio, "numel") else len(audio))
        io = _lib.SchedulerIO(0, 0, ptr(strength), ptr(dem), ptr(rec), ptr(audio), cap, 0)
        check(self.ctx.lib.rfa_scheduler_process(self.handle, ptr(packets), int(npackets), C.byref(io),
                                                 _mem_of(packets, audio)))
        return {"frames": io.frames, "signal_strength": strength[: io.frames].copy(), "demod_gate": dem, "record_gate": rec,
                "n_audio": io.n_audio}

    def state(self):
        """Device addresses of ring / peaks / avg and the counters (rfa_scheduler_state)."""
        ring, peaks, avg = C.c_void_p(), C.c_void_p(), C.c_void_p()
        newest, valid, pk, fr = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_longlong()
        sq, db = C.c_int(), C.c_int()
        check(self.ctx.lib.rfa_scheduler_state(self.handle, C.byref(ring), C.byref(newest), C.byref(valid), C.byref(peaks),
                                               C.byref(avg), C.byref(sq), C.byref(db), C.byref(pk), C.byref(fr)))
        return {"ring": ring.value, "newest_row": newest.value, "valid_rows": valid.value, "peaks": peaks.value,
                "avg": avg.value, "squelchSatisfied": bool(sq.value), "squelchDebounceCounter": db.value,
                "packets": pk.value, "frames": fr.value}

    def copy_state(self):
        """(ring[ring_rows][fft_size], peaks, avg) as numpy arrays (rfa_scheduler_read)."""
        n, rows = self.desc.fft_size, self.desc.ring_rows
        ring, peaks, avg = np.empty((rows, n), np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        check(self.ctx.lib.rfa_scheduler_read(self.handle, ptr(ring), ptr(peaks), ptr(avg)))
        return ring, peaks, avg

    def close(self):
        if self.handle:
            self.ctx.lib.rfa_scheduler_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IqConverterInt16:
    """iqconverter_int16 (libairspy/.../iqconverter_int16.c:54-208): Airspy / HydraSDR real ADC samples -> int16 IQ,
    in place, state carried across calls.  `samples`: int16 numpy array or CUDA tensor (viewed as int16)."""

    def __init__(self, ctx, hb_kernel):
        self.ctx = ctx
        k = np.ascontiguousarray(hb_kernel, np.int16)
        self.handle = C.c_void_p()
        check(ctx.lib.rfa_iqconverter_create(ctx.handle, ptr(k), len(k), C.byref(self.handle)))

    def reset(self):
        check(self.ctx.lib.rfa_iqconverter_reset(self.handle))

    def process(self, samples, length=None):
        from .engine import _mem_of
        n = int(length if length is not None else (samples.numel() if hasattr(samples, "numel") else len(samples)))
        check(self.ctx.lib.rfa_iqconverter_process(self.handle, ptr(samples), n, _mem_of(samples)))
        return samples

    def stats(self):
        """(chunks, chunks re-run sequentially, chunks skipped in the DC blocker's dead zone) since the last reset."""
        v = [C.c_longlong() for _ in range(3)]
        check(self.ctx.lib.rfa_iqconverter_stats(self.handle, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def close(self):
        if self.handle:
            self.ctx.lib.rfa_iqconverter_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def airspy_convert_samples(ctx, src, dst):
    """airspy.c:299-309: raw 12-bit ADC words (uint16) -> (raw - 2048) << 4 (int16)."""
    from .engine import _mem_of
    n = src.numel() if hasattr(src, "numel") else len(src)
    check(ctx.lib.rfa_airspy_convert_samples(ctx.handle, ptr(src), ptr(dst), int(n), _mem_of(src, dst)))


# ---- recordings on disk ---------------------------------------------------------------
FILE_HACKRF, FILE_RTLSDR, FILE_AIRSPY, FILE_HYDRASDR = range(4)   # FilesourceFileFormat


def parse_recording_name(ctx_or_lib, filename, file_format=FILE_HACKRF, frequency=0, sample_rate=0):
    """MainViewModel.setFilesourceUri (MainViewModel.kt:2034-2080): (file_format, frequency, sample_rate) with
    whatever the name carries replacing the values passed in."""
    lib = getattr(ctx_or_lib, "lib", ctx_or_lib)
    info = _lib.RecordingInfo(int(file_format), int(frequency), int(sample_rate), 0, 0, 0)
    check(lib.rfa_recording_parse_name(filename.encode(), C.byref(info)))
    return info.file_format, info.frequency, info.sample_rate


def recording_file_name(ctx_or_lib, timestamp, name, file_format, frequency, sample_rate):
    """Recording.calculateFileName (RecordingDao.kt:87-90)."""
    lib = getattr(ctx_or_lib, "lib", ctx_or_lib)
    buf = C.create_string_buffer(1024)
    check(lib.rfa_recording_file_name(timestamp.encode(), name.encode(), int(file_format), int(frequency),
                                      int(sample_rate), buf, len(buf)))
    return buf.value.decode()


class FileIQSource:
    """FileIQSource.java:41-395 (the parts a headless run needs): init / getPacket / getBytesPerSample."""
    FILE_FORMAT_8BIT_SIGNED, FILE_FORMAT_8BIT_UNSIGNED, FILE_FORMAT_16BIT_SIGNED = 0, 1, 2

    def __init__(self, ctx_or_lib, path, sampleRate, frequency, packetSize, repeat, fileFormat, pace=False):
        self.lib = getattr(ctx_or_lib, "lib", ctx_or_lib)
        self.sampleRate, self.frequency, self.packetSize, self.fileFormat = sampleRate, frequency, packetSize, fileFormat
        file_format = {0: FILE_HACKRF, 1: FILE_RTLSDR, 2: FILE_AIRSPY}[fileFormat]
        self.handle = C.c_void_p()
        check(self.lib.rfa_file_source_open(path.encode(), file_format, int(packetSize), 1 if repeat else 0,
                                            int(sampleRate) if pace else 0, C.byref(self.handle)))
        self.buffer = np.empty(packetSize, np.uint8)

    def getBytesPerSample(self):
        return 4 if self.fileFormat == self.FILE_FORMAT_16BIT_SIGNED else 2

    def getPacketSize(self):
        return self.packetSize

    def getPacket(self, timeout=0):
        """The packet buffer (reused between calls, like the reference's) or None at the end of the file."""
        rc = self.lib.rfa_file_source_get_packet(self.handle, self.buffer.ctypes.data)
        if rc < 0:
            check(-rc)
        return self.buffer if rc == 1 else None

    def close(self):
        if self.handle:
            self.lib.rfa_file_source_close(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
