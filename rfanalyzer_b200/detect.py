"""Signal detectors on the newest waterfall row -- host-side mirror of the reference's scan detectors
(ui/MainViewModel.kt), names and argument meaning kept:

  getAverageSignalLevel            :1392-1414      detectSignal              :1416-1461
  detectSignalsInFFT               :1463-1550      groupSignals              :1552-1607
  detectIEMChannelsInFFT           :861-935        detectAirCommSignal       :1151-1190
  detectAirCommSignalAtFrequency   :1202-1250      squelchSatisfied          database/AppStateRepository.kt:318-323

The window arithmetic (bin of a frequency, window half width, scan grid, grouping, detection mode) and the
reductions (peak / average per window, rfa_detect_windows: one launch for any number of windows, rows stay in
HBM) are the C ABI's; nothing is computed here.  `fftProcessorData` is dsp.FftProcessorData (device ring).
Batched entry points (`*_rows`) scan many rows per launch -- what a long recording needs.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import DETECT_PEAK_ONLY, DETECT_AVERAGE_ONLY, DETECT_PEAK_OR_AVERAGE, check

PEAK_ONLY, AVERAGE_ONLY, PEAK_OR_AVERAGE = DETECT_PEAK_ONLY, DETECT_AVERAGE_ONLY, DETECT_PEAK_OR_AVERAGE


@dataclass
class DiscoveredSignal:
    frequency: int
    peakStrength: float
    averageStrength: float
    bandwidth: int = 0
    isGrouped: bool = False


@dataclass
class IEMDetectedChannel:
    channelId: int
    peakStrength: float
    averageStrength: float
    detectionConfidence: float = 1.0


def _current_row(d):
    """The reference's guards: null / empty buffer, readIndex outside the ring -> no answer."""
    wb = d.waterfallBuffer
    if wb is None or wb.shape[0] == 0 or wb.shape[1] == 0:
        return None
    if d.readIndex < 0 or d.readIndex >= wb.shape[0]:
        return None
    return wb, int(d.readIndex), int(wb.shape[1])


def _lib_():
    return _lib.load()


def squelchSatisfied(averageSignalStrength, squelch, squelchEnabled):
    return averageSignalStrength > squelch if squelchEnabled else True


def getAverageSignalLevel(ctx, fftProcessorData):
    cur = _current_row(fftProcessorData)
    if cur is None:
        return None
    wb, row, n = cur
    _, avg = ctx.detect_windows(wb, wb.stride(0), n, [(row, 0, n - 1)])
    return float(avg[0])


def detectSignal(ctx, fftProcessorData, threshold, mode, noiseFloor, noiseFloorMargin):
    cur = _current_row(fftProcessorData)
    if cur is None:
        return None
    wb, row, n = cur
    peak, avg = ctx.detect_windows(wb, wb.stride(0), n, [(row, 0, n - 1)])
    hit = _lib_().rfa_detect_decide(float(peak[0]), float(avg[0]), float(threshold), float(noiseFloor),
                                    float(noiseFloorMargin), int(mode))
    if hit < 0:
        raise _lib.RfaError(_lib.ERR_INVALID if hasattr(_lib, "ERR_INVALID") else 1, _lib_().rfa_last_error().decode())
    return (float(peak[0]), float(avg[0])) if hit else None


def scan_grid(centerFrequency, sampleRate, usableBandwidth, stepSize, scanStartFreq, scanEndFreq, n, row=0):
    """Frequencies and +-2-bin windows of detectSignalsInFFT's while loop (rfa_scan_grid)."""
    lib = _lib_()
    args = (int(centerFrequency), int(sampleRate), int(usableBandwidth), int(stepSize), int(scanStartFreq),
            int(scanEndFreq), int(n), int(row))
    count = lib.rfa_scan_grid(*args, None, None, 0)
    if count < 0:
        raise _lib.RfaError(1, lib.rfa_last_error().decode())
    freqs = np.empty(count, np.int64)
    wins = (_lib.DetectWindow * max(count, 1))()
    lib.rfa_scan_grid(*args, freqs.ctypes.data, C.addressof(wins), count)
    return freqs, wins, count


def detectSignalsInFFT(ctx, fftProcessorData, centerFrequency, sampleRate, usableBandwidth, stepSize, threshold, mode,
                       noiseFloor, noiseFloorMargin, scanStartFreq, scanEndFreq):
    cur = _current_row(fftProcessorData)
    if cur is None:
        return []
    wb, row, n = cur
    return detectSignalsInFFT_rows(ctx, wb, [row], centerFrequency, sampleRate, usableBandwidth, stepSize, threshold,
                                   mode, noiseFloor, noiseFloorMargin, scanStartFreq, scanEndFreq)[0]


def detectSignalsInFFT_rows(ctx, rows, row_indices, centerFrequency, sampleRate, usableBandwidth, stepSize, threshold,
                            mode, noiseFloor, noiseFloorMargin, scanStartFreq, scanEndFreq):
    """detectSignalsInFFT for every row of `row_indices` in ONE launch; returns a list of signal lists."""
    n = int(rows.shape[1])
    freqs, wins, count = scan_grid(centerFrequency, sampleRate, usableBandwidth, stepSize, scanStartFreq, scanEndFreq, n)
    if count == 0:
        return [[] for _ in row_indices]
    allw = (_lib.DetectWindow * (count * len(row_indices)))()
    for k, r in enumerate(row_indices):
        for i in range(count):
            allw[k * count + i] = _lib.DetectWindow(int(r), wins[i].start, wins[i].end)
    peak, avg = ctx.detect_windows(rows, rows.stride(0), n, allw)
    lib = _lib_()
    out = []
    for k in range(len(row_indices)):
        sig = []
        for i in range(count):
            p, a = float(peak[k * count + i]), float(avg[k * count + i])
            if lib.rfa_detect_decide(p, a, float(threshold), float(noiseFloor), float(noiseFloorMargin), int(mode)) == 1:
                sig.append(DiscoveredSignal(int(freqs[i]), p, a, 0, False))
        out.append(sig)
    return out


def groupSignals(signals, stepSize, minimumGap):
    if not signals:
        return []
    n = len(signals)
    arr = (_lib.Signal * n)(*[_lib.Signal(int(s.frequency), float(s.peakStrength), float(s.averageStrength),
                                          int(s.bandwidth), int(s.isGrouped)) for s in signals])
    out = (_lib.Signal * n)()
    m = _lib_().rfa_group_signals(C.addressof(arr), n, int(stepSize), int(minimumGap), C.addressof(out))
    if m < 0:
        raise _lib.RfaError(1, _lib_().rfa_last_error().decode())
    return [DiscoveredSignal(o.frequency, o.peak, o.average, o.bandwidth, bool(o.grouped)) for o in out[:m]]


def _window_at(centerFrequency, sampleRate, n, freq, half_hz, min_half):
    b, s, e = C.c_int(), C.c_int(), C.c_int()
    inside = _lib_().rfa_detect_window_at(int(centerFrequency), int(sampleRate), int(n), int(freq), int(half_hz),
                                          int(min_half), C.byref(b), C.byref(s), C.byref(e))
    return inside == 1, b.value, s.value, e.value


def detectIEMChannelsInFFT(ctx, fftProcessorData, channels, centerFrequency, sampleRate, threshold):
    """`channels`: objects with .id and .frequency (database/IEMChannel)."""
    cur = _current_row(fftProcessorData)
    if cur is None:
        return []
    wb, row, n = cur
    picked, wins = [], []
    for ch in channels:
        inside, _, s, e = _window_at(centerFrequency, sampleRate, n, ch.frequency, 100000, 5)
        if inside:
            picked.append(ch)
            wins.append((row, s, e))
    if not wins:
        return []
    peak, avg = ctx.detect_windows(wb, wb.stride(0), n, wins)
    return [IEMDetectedChannel(ch.id, float(p), float(a), 1.0) for ch, p, a in zip(picked, peak, avg) if p > threshold]


def detectAirCommSignal(ctx, fftProcessorData, sampleRate):
    cur = _current_row(fftProcessorData)
    if cur is None:
        return None
    wb, row, n = cur
    half = _lib_().rfa_detect_half_width(int(sampleRate), n, 12500, 3)
    c = n // 2
    peak, _ = ctx.detect_windows(wb, wb.stride(0), n, [(row, max(c - half, 0), min(c + half, n - 1))])
    return float(peak[0])


def detectAirCommSignalAtFrequency(ctx, fftProcessorData, targetFreq, batchCenter, sampleRate):
    cur = _current_row(fftProcessorData)
    if cur is None:
        return None
    wb, row, n = cur
    inside, _, s, e = _window_at(batchCenter, sampleRate, n, targetFreq, 12500, 3)
    if not inside:
        return None
    peak, _ = ctx.detect_windows(wb, wb.stride(0), n, [(row, s, e)])
    return float(peak[0])
