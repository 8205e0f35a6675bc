"""rfa_scheduler_* (csrc/scheduler.cu) against the restated Scheduler loop (oracle/scheduler.py): FFT frame formation
from packets (first fftSize samples of a packet / several packets per frame, the rest dropped, Scheduler.kt:254-276),
ring / peak hold / average, per-frame channel strength, the squelch debounce of 50 iterations and the two gates
(:161-165, :199, :237), the audio of exactly the delivered packets; state carried across calls; host and device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DB_TOL = 0.01


def _signal(oracle, rfa, fmt, fs, n, burst):
    """A carrier 1/10 of the rate above the centre on a 24 dB higher noise floor inside `burst` = (first, last) sample; faint noise elsewhere."""
    off = fs // 10
    mul = 256 if fmt == 2 else 1
    quiet = oracle.synth_iq(fmt, n, comps=[], noise_shift=6)
    loud = oracle.synth_iq(fmt, n, comps=[(rfa.synth_step(off / fs), 60 * mul, rfa.synth_step(1000 / fs), 0),
                                          (rfa.synth_step((off + 1200) / fs), 20 * mul, 0, 0)], noise_shift=2)
    bps = 2 if fmt < 2 else 4
    out = quiet.copy()
    out[burst[0] * bps:burst[1] * bps] = loud[burst[0] * bps:burst[1] * bps]
    return out, 100_000_000, 100_000_000 + off


def _rows_close(got, want):
    """0.01 dB, or -- for the few bins of a 2.6-million-bin ring that are deep random nulls, where a float32 FFT's own
    rounding is the value -- BASELINE's linear-power bound with its floor of 1e-6 of the row's strongest bin."""
    ok = np.abs(got - want) < DB_TOL
    lin_g, lin_w = 10.0 ** (got.astype(np.float64) / 5.0), 10.0 ** (want.astype(np.float64) / 5.0)
    ok |= np.abs(lin_g - lin_w) <= 1e-4 * lin_w + 1e-6 * lin_w.max(axis=-1, keepdims=True)
    return bool(ok.all())


CASES = [
    # fmt, fs, packet, fft, mode, width, squelch_enabled, rec_only, npackets
    (1, 2_400_000, 8192, 4096, 3, 100_000, True, True, 150),     # RTL-SDR: frame = first half of every packet, wFM, squelch
    (0, 2_000_000, 16384, 1024, 1, 8_000, True, False, 140),     # frame = first 1/16 of a packet, AM
    (2, 10_000_000, 65536, 65536, 2, 10_000, True, True, 130),   # Airspy: packet == frame, nFM
    (1, 2_400_000, 8192, 32768, 3, 100_000, False, False, 30),   # four packets per frame, squelch off: chain beside spectrum
    (0, 2_000_000, 4096, 16384, 0, 0, True, True, 70),           # no demodulator: strength never updates
]


@pytest.mark.parametrize("fmt,fs,packet,fft,mode,width,sq_on,rec_only,npk", CASES)
@pytest.mark.parametrize("device", [True, False])
def test_scheduler_vs_restated_loop(gpu_ctx, oracle, fmt, fs, packet, fft, mode, width, sq_on, rec_only, npk, device):
    import torch
    import rfanalyzer_b200 as rfa
    from oracle import scheduler as OS
    n = packet * npk
    iq, src, chan = _signal(oracle, rfa, fmt, fs, n, (packet * 20, packet * 45))
    bps = 2 if fmt < 2 else 4
    L, ring_rows = 3, 40
    # squelch threshold between the quiet and the loud channel strength
    probe = OS.scheduler_run(fmt, iq, packet, fs, src, fft, L, ring_rows, mode or 2, chan, width, squelch_enabled=False)
    st = probe["strengths"]
    squelch = float((st.min() + st.max()) / 2) if len(st) and mode else -30.0
    sched = rfa.Scheduler(gpu_ctx, fmt, fs, src, packet, fft, avg_len=L, ring_rows=ring_rows, mode=mode, channelFrequency=chan,
                          channelWidth=width, volume=0.8, flags=rfa.SUM_EXACT, squelchEnabled=sq_on, squelch=squelch,
                          recordOnlyWhenSquelchIsSatisfied=rec_only)
    state, got_audio, gates_d, gates_r, strengths = None, [], [], [], []
    cuts = sorted({min(c, npk) for c in (0, 7, 8, 70, npk)})   # several calls: partial frames and counters carry over
    for a, b in zip(cuts, cuts[1:]):
        part = iq[a * packet * bps:b * packet * bps]
        want = OS.scheduler_run(fmt, part, packet, fs, src, fft, L, ring_rows, mode, chan, width, 0.8, sq_on, squelch, rec_only, state)
        state = want["state"]
        if device:
            with torch.cuda.stream(gpu_ctx.torch_stream):
                d = torch.from_numpy(part.copy()).cuda()
                audio = torch.zeros(sched.max_audio(b - a), dtype=torch.float32, device="cuda")
                gpu_ctx.sync()
            res = sched.process(d, b - a, audio if mode else None)
            if mode:
                got_audio.append(audio[: res["n_audio"]].cpu().numpy())
        else:
            audio = np.zeros(sched.max_audio(b - a), np.float32)
            res = sched.process(part.copy(), b - a, audio if mode else None)
            if mode:
                got_audio.append(audio[: res["n_audio"]].copy())
        assert res["frames"] == len(want["strengths"])
        assert np.array_equal(res["demod_gate"], want["demod_gate"]), (a, b)
        assert np.array_equal(res["record_gate"], want["record_gate"]), (a, b)
        if mode:
            assert np.abs(res["signal_strength"] - want["strengths"]).max() < DB_TOL
        ring, peaks, avg = sched.copy_state()
        assert _rows_close(ring, want["ring"])
        assert _rows_close(peaks, want["peaks"])
        fin = np.abs(want["avg"]) < 9000
        assert np.abs(avg - want["avg"])[fin].max() < DB_TOL if fin.any() else True
        gates_d.append(res["demod_gate"])
    gates_d = np.concatenate(gates_d)
    if mode and sq_on:
        # the scenario exercises what it claims: the gate opens with the burst and closes 50 iterations after it
        assert gates_d[:20].all() and not gates_d.all() and gates_d[25:45].all()
        first_closed = int(np.argmin(gates_d))
        assert first_closed > 45
    if mode:
        want_audio = OS.delivered_audio(state, fmt, fs, src, chan, mode, width, packet, 0.8)
        got = np.concatenate(got_audio)
        assert len(got) == len(want_audio)
        if mode in (2, 3):
            assert np.abs(got - want_audio).max() <= 1e-6 * np.abs(want_audio).max()
        else:
            assert np.array_equal(got, want_audio)
    st = sched.state()
    assert st["packets"] == npk and st["squelchSatisfied"] == state["squelch"] and st["squelchDebounceCounter"] == state["counter"]
