#!/usr/bin/env python
"""Throughput of the IQ->audio chain (BASELINE configs 2 and 4) on one B200, device-resident input,
beside the CPU oracle (the line-faithful port of the reference's JVM chain) on a bounded sample.
Secondary to bench.py (the headline metric is the FFT+waterfall path); writes one JSON line per config."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import rfanalyzer_b200 as rfa

CONFIGS = [
    ("C2 RTL-SDR uint8 @2.4 Msps -> wFM", rfa.FMT_U8, 2_400_000, rfa.MODE_WFM, 100_000, 8192),
    ("C4 Airspy int16 @10 Msps -> nFM", rfa.FMT_S16LE, 10_000_000, rfa.MODE_NFM, 10_000, 65536),
    ("C4 Airspy int16 @10 Msps -> USB", rfa.FMT_S16LE, 10_000_000, rfa.MODE_USB, 2_800, 65536),
    ("C4 Airspy int16 @10 Msps -> CW", rfa.FMT_S16LE, 10_000_000, rfa.MODE_CW, 300, 65536),
    ("C4 Airspy int16 @10 Msps -> AM", rfa.FMT_S16LE, 10_000_000, rfa.MODE_AM, 8_000, 65536),
]


def main():
    S = 1 << int(os.environ.get("LOG2_SAMPLES", "26"))
    with_cpu = os.environ.get("CPU", "1") == "1"
    stream = torch.cuda.Stream()
    ctx = rfa.Context(0, stream)
    for name, fmt, fs, mode, width, packet in CONFIGS:
        off = fs // 10
        mul = 256 if fmt == rfa.FMT_S16LE else 1
        comps = [(rfa.synth_step(off / fs), 60 * mul, rfa.synth_step(1000 / fs), 3_130_000 if mode in (2, 3) else 0),
                 (rfa.synth_step((off + 1200) / fs), 20 * mul, 0, 0)]
        bps = rfa.BYTES_PER_SAMPLE[fmt]
        with torch.cuda.stream(stream):
            iq = torch.empty(S * bps, dtype=torch.uint8, device="cuda")
            rfa.synth_iq(ctx, fmt, S, iq, comps=comps, noise_shift=3)
            out = {}
            for flags, label in ((rfa.SUM_FMA, "fma"), (rfa.SUM_EXACT, "exact")):
                plan = rfa.ChainPlan(ctx, fmt, fs, 100_000_000, 100_000_000 + off, mode, width, packet, 1.0, flags)
                audio = torch.empty(plan.max_audio(S), dtype=torch.float32, device="cuda")
                plan.process(iq, S, audio)
                stream.synchronize()
                reps = 5
                l0 = ctx.launch_count
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
                for _ in range(reps):
                    n_audio = plan.process(iq, S, audio)
                ev1.record(stream)
                stream.synchronize()
                ms = ev0.elapsed_time(ev1) / reps
                alg_bytes = S * bps + 4 * n_audio
                out[label] = {"Msamples/s": S / ms / 1e3, "ms": ms, "GB/s_algorithmic": alg_bytes / ms / 1e6,
                              "launches_per_call": (ctx.launch_count - l0) / reps,
                              "I/D": [plan.interpolation, plan.decimation], "taps_per_phase": plan.taps_per_phase}
                plan.close()
        line = {"config": name, "samples": S, "gpu": out}
        if with_cpu:
            from oracle import oracle as O
            ns = 1 << 22
            h = O.synth_iq(fmt, ns, comps=comps, noise_shift=3)
            t0 = time.perf_counter()
            O.chain_run(fmt, h, fs, 100_000_000, 100_000_000 + off, mode, width, packet)
            dt = time.perf_counter() - t0
            line["cpu_port_1_thread"] = {"Msamples/s": ns / dt / 1e6, "sample": "2^22 samples, %.1f s" % dt}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
