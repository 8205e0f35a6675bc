"""TEST INFRASTRUCTURE (oracle): restatement of the reference's scan detectors, line by line, in numpy float32 /
Python float (= Java double) arithmetic.  Only tests/ may import this.  Parity unpinned by reference tests (the
reference has none for these functions); anchored on the source lines cited below.

  ui/MainViewModel.kt  getAverageSignalLevel :1392-1414, detectSignal :1416-1461, detectSignalsInFFT :1463-1550,
                       groupSignals/finalizeGroup :1552-1607, detectIEMChannelsInFFT :861-935,
                       detectAirCommSignal :1151-1190, detectAirCommSignalAtFrequency :1202-1250
"""
import math

import numpy as np

F = np.float32
PEAK_ONLY, AVERAGE_ONLY, PEAK_OR_AVERAGE = 0, 1, 2


def to_int(v):
    """Float.toInt(): truncate toward zero, saturate, NaN -> 0."""
    v = float(v)
    if math.isnan(v):
        return 0
    if v >= 2147483648.0:
        return 2147483647
    if v <= -2147483648.0:
        return -2147483648
    return int(v)


def jdiv(a, b):
    """Java integer division: truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def jmax(a, b):
    """Math.max(float, float): NaN if either is NaN."""
    if math.isnan(a) or math.isnan(b):
        return F(np.nan)
    return a if a >= b else b


def max_or_null(w):
    if len(w) == 0:
        return None
    m = w[0]
    for e in w[1:]:
        m = jmax(m, e)
    return F(m)


def average(w):
    """FloatArray.average(): Double sum in index order / count."""
    s = 0.0
    for e in w:
        s += float(e)
    return s / len(w) if len(w) else float("nan")


def resolution(sampleRate, fftSize):
    return F(F(sampleRate) / F(fftSize))     # sampleRate.toFloat() / fftSize


def bin_index(freq, startFrequency, res):
    return to_int(F(F(freq - startFrequency) / res))   # Long / Float -> Float


def decide(peak, avg, threshold, noiseFloor, margin, mode):
    eff = max(F(threshold), F(F(noiseFloor) + F(margin)))
    if mode == PEAK_ONLY:
        return bool(peak > eff)
    if mode == AVERAGE_ONLY:
        return bool(avg > eff)
    return bool(peak > eff or avg > eff)


def get_average_signal_level(row):
    return F(average(row))


def detect_signal(row, threshold, mode, noiseFloor, margin):
    peak, avg = max_or_null(row), F(average(row))
    return (peak, avg) if decide(peak, avg, threshold, noiseFloor, margin, mode) else None


def detect_signals_in_fft(row, centerFrequency, sampleRate, usableBandwidth, stepSize, threshold, mode, noiseFloor,
                          margin, scanStartFreq, scanEndFreq):
    n = len(row)
    res = resolution(sampleRate, n)
    startFrequency = centerFrequency - jdiv(sampleRate, 2)
    usableStartOffset = jdiv(sampleRate - usableBandwidth, 2)
    usableEndOffset = usableStartOffset + usableBandwidth
    out = []
    f = max(scanStartFreq, startFrequency + usableStartOffset)
    endFreq = min(scanEndFreq, startFrequency + usableEndOffset)
    while f <= endFreq:
        b = bin_index(f, startFrequency, res)
        if 0 <= b < n:
            w = row[max(0, b - 2):min(n - 1, b + 2) + 1]
            peak, avg = max_or_null(w), F(average(w))
            if decide(peak, avg, threshold, noiseFloor, margin, mode):
                out.append((f, peak, avg, 0, False))
        f += stepSize
    return out


def group_signals(signals, stepSize, minimumGap):
    """signals: (frequency, peak, avg, bandwidth, grouped) tuples."""
    if not signals:
        return []
    s = sorted(signals, key=lambda t: t[0])
    gap = stepSize * minimumGap
    groups, cur = [], [s[0]]
    for prev, now in zip(s, s[1:]):
        if now[0] - prev[0] <= gap:
            cur.append(now)
        else:
            groups.append(cur)
            cur = [now]
    groups.append(cur)
    out = []
    for g in groups:
        if len(g) == 1:
            out.append(g[0])
            continue
        mn, mx = min(t[0] for t in g), max(t[0] for t in g)
        pk = g[0][1]
        for t in g[1:]:
            pk = jmax(pk, t[1])
        out.append((jdiv(mn + mx, 2), F(pk), F(average([t[2] for t in g])), mx - mn, True))
    return out


def detect_iem_channels(row, channels, centerFrequency, sampleRate, threshold):
    """channels: (id, frequency) pairs -> (id, peak, avg) of the detected ones."""
    n = len(row)
    res = resolution(sampleRate, n)
    startFrequency = centerFrequency - jdiv(sampleRate, 2)
    out = []
    for cid, freq in channels:
        b = bin_index(freq, startFrequency, res)
        if 0 <= b < n:
            half = max(to_int(F(F(100000) / res)), 5)
            w = row[max(b - half, 0):min(b + half, n - 1) + 1]
            peak, avg = max_or_null(w), F(average(w))
            if peak > F(threshold):
                out.append((cid, peak, avg))
    return out


def detect_aircomm_signal(row, sampleRate):
    n = len(row)
    res = resolution(sampleRate, n)
    half = max(to_int(F(F(12500) / res)), 3)
    c = n // 2
    return max_or_null(row[max(c - half, 0):min(c + half, n - 1) + 1])


def detect_aircomm_signal_at_frequency(row, targetFreq, batchCenter, sampleRate):
    n = len(row)
    res = resolution(sampleRate, n)
    startFrequency = batchCenter - jdiv(sampleRate, 2)
    b = bin_index(targetFreq, startFrequency, res)
    if not 0 <= b < n:
        return None
    half = max(to_int(F(F(12500) / res)), 3)
    return max_or_null(row[max(b - half, 0):min(b + half, n - 1) + 1])
