"""Timing sweep on the GPU box (device-resident inputs, rotating buffers > L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rfanalyzer_b200 as rfa
from oracle import oracle as O

stream = torch.cuda.Stream()
ctx = rfa.Context(0, stream)
for kv in filter(None, os.environ.get("KNOBS", "").split(",")):  # e.g. KNOBS=cluster=0,pdl=0
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
reps, nbuf = 30, 6
fmt = int(os.environ.get("FMT", "0"))
sizes = [int(x) for x in os.environ.get("SIZES", "4096,1024,2048,8192,16384,32768,65536").split(",")]
variants = os.environ.get("VARIANTS", "rows+peaks+avg,rows,rows+peaks,rows+avg,peaks").split(",")
with torch.cuda.stream(stream):
    for N in sizes:
        F = (1 << 24) // N
        plan = rfa.SpectrumPlan(ctx, fmt, N, avg_len=8)
        base = torch.from_numpy(O.synth_iq(fmt, N * F)).cuda()
        iqs = [base.clone() for i in range(nbuf)]
        rows = [torch.empty((F, N), dtype=torch.float32, device="cuda") for _ in range(nbuf)]
        peaks = torch.zeros(N, dtype=torch.float32, device="cuda"); avg = torch.zeros(N, dtype=torch.float32, device="cuda")
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        for var in variants:
            kw = lambda i: dict(rows=rows[i % nbuf] if "rows" in var else None, peaks=peaks if "peaks" in var else None,
                                avg=avg if "avg" in var else None, peaks_accumulate=True)
            for i in range(nbuf):
                plan.process(iqs[i], F, **kw(i))
            stream.synchronize()
            ev0.record(stream)
            for i in range(reps):
                plan.process(iqs[i % nbuf], F, **kw(i))
            ev1.record(stream); stream.synchronize()
            ms = ev0.elapsed_time(ev1) / reps
            b = plan.algorithmic_bytes(F, "rows" in var)
            print(f"fmt={fmt} N={N} F={F} [{var}]: {ms*1e3:.1f} us/call  {N*F/ms/1e3:.1f} Msamples/s  "
                  f"{b/ms/1e6:.1f} GB/s  frac={b/ms/1e6/6531.9:.3f}", flush=True)
        del iqs, rows, plan
