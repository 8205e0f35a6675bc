"""Small pass over the main device paths (every spectrum kernel family, detectors, one demodulation chain): a quick
all-paths run for a fresh box, and a target for compute-sanitizer where the pool allows it (this one does not)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rfanalyzer_b200 as rfa
from oracle import oracle as O

stream = torch.cuda.Stream()
ctx = rfa.Context(0, stream)
with torch.cuda.stream(stream):
    for fmt, n, frames in ((0, 4096, 9), (1, 1024, 40), (2, 256, 70), (0, 8192, 5), (0, 16384, 3), (0, 65536, 3), (2, 32768, 3)):
        plan = rfa.SpectrumPlan(ctx, fmt, n, avg_len=3)
        iq = torch.from_numpy(O.synth_iq(fmt, n * frames)).cuda()
        rows = torch.zeros((frames, n), dtype=torch.float32, device="cuda")
        peaks = torch.zeros(n, dtype=torch.float32, device="cuda"); avg = torch.zeros(n, dtype=torch.float32, device="cuda")
        for rep in range(2):
            plan.process(iq, frames, rows=rows, peaks=peaks, avg=avg, peaks_accumulate=rep > 0)
        ctx.sync()
        peak, mean = ctx.detect_windows(rows, n, n, [(0, 0, n - 1), (frames - 1, 5, 50)])
        print(fmt, n, float(rows.max()), float(peak[0]))
    for var, kernel in (("pair", 3), ("lean", 4)):   # lab build only (RFA_B200_LIB=.../librfa_b200_lab.so)
        try:
            ctx.set_option("kernel", kernel)
        except rfa.RfaError:
            print(var, "kernel: not in this build")
            continue
        plan = rfa.SpectrumPlan(ctx, 0, 4096, avg_len=3)
        iq = torch.from_numpy(O.synth_iq(0, 4096 * 11)).cuda()
        rows = torch.zeros((11, 4096), dtype=torch.float32, device="cuda")
        peaks = torch.zeros(4096, dtype=torch.float32, device="cuda"); avg = torch.zeros(4096, dtype=torch.float32, device="cuda")
        plan.process(iq, 11, rows=rows, peaks=peaks, avg=avg)
        ctx.sync()
        ctx.set_option("kernel", 0)
        print(var, float(rows.max()))
    fs = 2_400_000
    chain = rfa.ChainPlan(ctx, rfa.FMT_U8, fs, 100_000_000, 100_250_000, rfa.MODE_WFM, 100_000, 8192, 1.0, rfa.SUM_FMA)
    S = 1 << 18
    iq = torch.empty(S * 2, dtype=torch.uint8, device="cuda")
    rfa.synth_iq(ctx, rfa.FMT_U8, S, iq)
    audio = torch.empty(chain.max_audio(S), dtype=torch.float32, device="cuda")
    print("audio", chain.process(iq, S, audio))
    ctx.sync()
print("done")
