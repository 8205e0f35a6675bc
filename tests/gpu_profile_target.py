"""Small fixed workload for ncu: C1 shape (int8, N=4096, 4096 frames), a few launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rfanalyzer_b200 as rfa
from oracle import oracle as O

N = int(os.environ.get("N", "4096")); fmt = int(os.environ.get("FMT", "0"))
F = (1 << 24) // N
stream = torch.cuda.Stream()
ctx = rfa.Context(0, stream)
plan = rfa.SpectrumPlan(ctx, fmt, N, avg_len=8)
with torch.cuda.stream(stream):
    iq = torch.from_numpy(O.synth_iq(fmt, N * F)).cuda()
    iqs = [iq.clone() for _ in range(4)]
    rows = [torch.empty((F, N), dtype=torch.float32, device="cuda") for _ in range(4)]
    peaks = torch.zeros(N, dtype=torch.float32, device="cuda"); avg = torch.zeros(N, dtype=torch.float32, device="cuda")
    var = os.environ.get("VARIANT", "rows+peaks+avg")
    for i in range(8):
        plan.process(iqs[i % 4], F, rows=rows[i % 4] if "rows" in var else None, peaks=peaks if "peaks" in var else None,
                     avg=avg if "avg" in var else None, peaks_accumulate=i > 0)
    stream.synchronize()
print("done")
