#!/usr/bin/env python
"""Steady-state DRAM traffic of the fused spectrum kernel: target program for an ncu capture.

    ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        -k regex:spectrum_kernel --csv --log-file profiles/r02_traffic_launches.csv python tools/traffic_capture.py
    ncu --replay-mode range ... (same metrics): ONE measurement over the bracketed range of launches

BASELINE config 1 launches (int8, 2^24 samples, N = 4096, rows + peak hold + avg = 8) over EIGHT rotating input / output
sets (768 MiB, six times the L2), warmed up first, so that inside the window of launch k the rows of launch k-1 are
written back: per-launch dram bytes then read ~ algorithmic bytes (100.7 MB) instead of the cold single-launch
figure, whose 64 MiB of rows were still sitting in the 126 MB L2 when the window closed.
tools/traffic_summary.py turns the csv into profiles/traffic.json."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rfanalyzer_b200 as rfa

N = int(os.environ.get("N", "4096"))
FMT = int(os.environ.get("FMT", "0"))
LAUNCHES = int(os.environ.get("LAUNCHES", "16"))
S = 1 << 24
F = S // N
NBUF = 8
stream = torch.cuda.Stream()
ctx = rfa.Context(0, stream)
for kv in filter(None, os.environ.get("KNOBS", "").split(",")):  # e.g. KNOBS=cluster=1 with N=65536
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
plan = rfa.SpectrumPlan(ctx, FMT, N, avg_len=8)
bps = rfa.BYTES_PER_SAMPLE[FMT]
with torch.cuda.stream(stream):
    iqs, rows = [], []
    for j in range(NBUF):
        b = torch.empty(S * bps, dtype=torch.uint8, device="cuda")
        rfa.synth_iq(ctx, FMT, S, b, first=j * S)
        iqs.append(b)
        rows.append(torch.empty((F, N), dtype=torch.float32, device="cuda"))
    peaks = torch.full((N,), -999999.0, dtype=torch.float32, device="cuda")
    avg = torch.zeros(N, dtype=torch.float32, device="cuda")
    for k in range(2 * NBUF):                      # warm-up: L2 holds dirty rows of earlier launches, as in steady state
        plan.process(iqs[k % NBUF], F, rows=rows[k % NBUF], peaks=peaks, avg=avg, peaks_accumulate=True)
    stream.synchronize()
    torch.cuda.profiler.start()
    for k in range(LAUNCHES):
        plan.process(iqs[k % NBUF], F, rows=rows[k % NBUF], peaks=peaks, avg=avg, peaks_accumulate=True)
    stream.synchronize()
    torch.cuda.profiler.stop()
print("done: %d launches, algorithmic bytes per launch %d" % (LAUNCHES, plan.algorithmic_bytes(F, True)))
