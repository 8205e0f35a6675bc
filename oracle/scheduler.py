"""Scheduler.run restated (TEST INFRASTRUCTURE ONLY) -- A/analyzer/Scheduler.kt:140-298 in its loss-free, synchronous
reading: one loop iteration per packet, buffer pools never empty, and the FftProcessor (FftProcessor.kt:135-157) has
finished a delivered frame -- row, peak hold, averageSignalStrength -> squelchSatisfied (AppStateRepository.kt:318-323) --
before the next packet is taken.  Built from the oracle's own converter / FFT / FftProcessor / chain restatements."""
import ctypes as C

import numpy as np

from . import oracle as O

SQUELCH_DEBOUNCE_COUNT = 50  # Scheduler.kt:52


def scheduler_run(fmt, packets, packet_samples, sample_rate, frequency, fft_size, avg_len, ring_rows, mode=0,
                  channel_frequency=0, channel_width=0, volume=1.0, squelch_enabled=False, squelch=-30.0,
                  record_only_when_squelch_satisfied=False, state=None):
    """-> dict(ring, peaks, avg, strengths, demod_gate, record_gate, audio, state).  `state` carries the scheduler
    across calls (pass the returned one back in)."""
    L = O.lib()
    bps = O.BYTES_PER_SAMPLE[fmt]
    pbytes = packet_samples * bps
    npk = len(packets) // pbytes
    if state is None:
        conv = L.orc_converter_new(fmt)
        L.orc_converter_set_frequency(conv, int(frequency))
        L.orc_converter_set_sample_rate(conv, int(sample_rate))
        width = 0
        if mode != 0:
            d = L.orc_demod_new(48000)
            L.orc_demod_set_mode(d, mode)
            if channel_width:
                L.orc_demod_set_channel_width(d, int(channel_width))
            width = L.orc_demod_channel_width(d)
            L.orc_demod_free(d)
        state = {"conv": conv, "proc": L.orc_fftproc_new(ring_rows, 1), "buf": O.PacketView(fft_size),
                 "squelch": not squelch_enabled, "counter": 0, "delivered": [], "width": width}
        state["buf"].size = 0
    conv, proc, buf = state["conv"], state["proc"], state["buf"]
    demod = mode != 0
    dem, rec, strengths = np.zeros(npk, np.uint8), np.zeros(npk, np.uint8), []
    mag = np.empty(fft_size, np.float32)
    for k in range(npk):
        pkt = np.ascontiguousarray(packets[k * pbytes:(k + 1) * pbytes])
        # :161-165
        if state["squelch"]:
            state["counter"] = 0
        elif state["counter"] < SQUELCH_DEBOUNCE_COUNT:
            state["counter"] += 1
        # :199
        rec[k] = state["squelch"] or (not record_only_when_squelch_satisfied) or state["counter"] < SQUELCH_DEBOUNCE_COUNT
        # :237-244
        if demod and (state["squelch"] or state["counter"] < SQUELCH_DEBOUNCE_COUNT):
            dem[k] = 1
            state["delivered"].append(pkt)
        # :254-276
        L.orc_converter_fill(conv, pkt, pbytes, buf.p)
        if buf.size == buf.capacity:
            ok = L.orc_windowed_fft_logmag(buf.re, buf.im, fft_size, fft_size, fft_size, mag)
            assert ok
            L.orc_fftproc_push(proc, mag, fft_size, int(frequency), int(sample_rate))
            if demod:  # AnalyzerService.kt:344-353: no demodulator, no channel range, no strength update
                out = C.c_float()
                if L.orc_signal_strength(mag, fft_size, int(frequency), int(sample_rate), int(channel_frequency - state["width"]),
                                         int(channel_frequency + state["width"]), C.byref(out)):
                    strengths.append(out.value)
                    if squelch_enabled:
                        state["squelch"] = out.value > squelch
                else:
                    strengths.append(-999.0)
            else:
                strengths.append(-999.0)
            buf.size = 0
    ring = np.stack([np.ctypeslib.as_array(L.orc_fftproc_row(proc, i), shape=(fft_size,)).copy() for i in range(ring_rows)])
    peaks = np.ctypeslib.as_array(L.orc_fftproc_peaks(proc), shape=(fft_size,)).copy()
    avg = np.empty(fft_size, np.float32)
    L.orc_time_average(proc, avg_len, avg)
    return {"ring": ring, "peaks": peaks, "avg": avg, "strengths": np.array(strengths, np.float32), "demod_gate": dem,
            "record_gate": rec, "state": state}


def delivered_audio(state, fmt, sample_rate, frequency, channel_frequency, mode, channel_width, packet_samples, volume=1.0):
    """The audio of every packet the scheduler delivered so far (the Demodulator sees them back to back)."""
    if not state["delivered"]:
        return np.zeros(0, np.float32)
    iq = np.concatenate(state["delivered"])
    return O.chain_run(fmt, iq, sample_rate, frequency, channel_frequency, mode, channel_width, packet_samples, volume)
