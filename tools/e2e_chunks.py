"""e2e throughput of the host-buffer spectrum path against the chunk size of its H2D -> kernel -> D2H pipeline (knob chunk_kib)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rfanalyzer_b200 as rfa
N, S = 4096, 1 << 24
F = S // N
stream = torch.cuda.Stream(); ctx = rfa.Context(0, stream)
plan = rfa.SpectrumPlan(ctx, rfa.FMT_S8, N, avg_len=8)
iq = torch.empty(S * 2, dtype=torch.uint8).pin_memory(); iq.random_(0, 255)
rows = torch.empty((F, N), dtype=torch.float32).pin_memory()
peaks = torch.empty(N, dtype=torch.float32).pin_memory(); avg = torch.empty(N, dtype=torch.float32).pin_memory()
for kib in [int(k) for k in os.environ.get("KIBS", "2048,4096,8192,16384,4096,8192").split(",")]:
    ctx.set_option("chunk_kib", kib)
    for _ in range(3):
        plan.process(iq.numpy(), F, rows=rows.numpy(), peaks=peaks.numpy(), avg=avg.numpy())
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        plan.process(iq.numpy(), F, rows=rows.numpy(), peaks=peaks.numpy(), avg=avg.numpy())
    dt = (time.perf_counter() - t0) / reps
    print("chunk_kib %6d: %.3f ms per 2^24 samples = %.2f Gsamples/s" % (kib, dt * 1e3, S / dt / 1e9), flush=True)
