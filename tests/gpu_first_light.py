"""First-light script run on the GPU box: fused spectrum path vs the oracle for several
sizes/formats, device and host mode, plus a quick timing.  (The pytest suite supersedes it.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rfanalyzer_b200 as rfa
from oracle import oracle as O

ctx = rfa.Context(0, torch.cuda.current_stream())
print("SMs", ctx.sm_count, flush=True)
ok = True
for fmt in (0, 1, 2):
    for N in (16, 64, 256, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        F = 5 if N >= 8192 else 37
        iq = O.synth_iq(fmt, N * F)
        r_ref, p_ref, a_ref = O.ref_spectrum_run(fmt, iq, N, 3) if N >= 64 else O.spectrum_run(fmt, iq, N, 3)
        plan = rfa.SpectrumPlan(ctx, fmt, N, avg_len=3)
        d_iq = torch.from_numpy(iq).cuda()
        rows = torch.zeros((F, N), dtype=torch.float32, device="cuda")
        peaks = torch.zeros(N, dtype=torch.float32, device="cuda")
        avg = torch.zeros(N, dtype=torch.float32, device="cuda")
        plan.process(d_iq, F, rows=rows, peaks=peaks, avg=avg)
        torch.cuda.synchronize()
        e_r = np.abs(rows.cpu().numpy() - r_ref).max()
        e_p = np.abs(peaks.cpu().numpy() - p_ref).max()
        e_a = np.abs(avg.cpu().numpy() - a_ref).max()
        # host mode
        h_rows = np.zeros((F, N), np.float32); h_p = np.zeros(N, np.float32); h_a = np.zeros(N, np.float32)
        plan.process(iq, F, rows=h_rows, peaks=h_p, avg=h_a)
        e_h = max(np.abs(h_rows - r_ref).max(), np.abs(h_p - p_ref).max(), np.abs(h_a - a_ref).max())
        flag = "OK" if max(e_r, e_p, e_a, e_h) < 0.01 else "FAIL"
        ok &= flag == "OK"
        print(f"fmt={fmt} N={N:6d} rows {e_r:.2e} peaks {e_p:.2e} avg {e_a:.2e} host {e_h:.2e} {flag}", flush=True)
print("ALL OK" if ok else "SOME FAILED", flush=True)

# timing, C1 shape rotated over buffers larger than L2
N = 4096
F = 4096
nbuf = 6
plan = rfa.SpectrumPlan(ctx, 0, N, avg_len=8)
iqs = [torch.from_numpy(O.synth_iq(0, N * F, first=i * N * F)).cuda() for i in range(nbuf)]
rows = [torch.empty((F, N), dtype=torch.float32, device="cuda") for _ in range(nbuf)]
peaks = torch.zeros(N, dtype=torch.float32, device="cuda"); avg = torch.zeros(N, dtype=torch.float32, device="cuda")
for i in range(nbuf):
    plan.process(iqs[i], F, rows=rows[i], peaks=peaks, avg=avg)
torch.cuda.synchronize()
ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
reps = 30
ev0.record()
for i in range(reps):
    plan.process(iqs[i % nbuf], F, rows=rows[i % nbuf], peaks=peaks, avg=avg, peaks_accumulate=True)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
bytes_ = plan.algorithmic_bytes(F)
print(f"N=4096 F=4096: {ms*1e3:.1f} us/call  {N*F/ms/1e3:.1f} Msamples/s  {bytes_/ms/1e6:.1f} GB/s", flush=True)
for N in (1024, 2048, 8192, 16384, 32768, 65536):
    F = (1 << 24) // N
    plan = rfa.SpectrumPlan(ctx, 0, N, avg_len=8)
    iqs = [torch.from_numpy(O.synth_iq(0, N * F, first=i * N * F)).cuda() for i in range(nbuf)]
    rows = [torch.empty((F, N), dtype=torch.float32, device="cuda") for _ in range(nbuf)]
    peaks = torch.zeros(N, dtype=torch.float32, device="cuda"); avg = torch.zeros(N, dtype=torch.float32, device="cuda")
    for i in range(nbuf):
        plan.process(iqs[i], F, rows=rows[i], peaks=peaks, avg=avg)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(reps):
        plan.process(iqs[i % nbuf], F, rows=rows[i % nbuf], peaks=peaks, avg=avg, peaks_accumulate=True)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"N={N} F={F}: {ms*1e3:.1f} us/call  {N*F/ms/1e3:.1f} Msamples/s  {plan.algorithmic_bytes(F)/ms/1e6:.1f} GB/s", flush=True)
