"""Segment timeline of the anti-phase pair kernel (-DRFA_TRACE -DRFA_LAB build, knob kernel=3): cycles per segment and barrier wait."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rfanalyzer_b200 as rfa
from rfanalyzer_b200 import _lib
from oracle import oracle as O

N = 4096; F = (1 << 24) // N
stream = torch.cuda.Stream(); ctx = rfa.Context(0, stream)
ctx.set_option("kernel", 3)
plan = rfa.SpectrumPlan(ctx, 0, N, avg_len=8)
with torch.cuda.stream(stream):
    iq = torch.from_numpy(O.synth_iq(0, N * F)).cuda()
    iqs = [iq.clone() for _ in range(6)]
    rows = [torch.empty((F, N), dtype=torch.float32, device="cuda") for _ in range(6)]
    peaks = torch.zeros(N, dtype=torch.float32, device="cuda"); avg = torch.zeros(N, dtype=torch.float32, device="cuda")
    for i in range(12):
        plan.process(iqs[i % 6], F, rows=rows[i % 6], peaks=peaks, avg=avg, peaks_accumulate=True)
    stream.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
ctas, slots = 147, 128
buf = np.zeros((ctas, slots), np.int64)
assert lib.rfa_debug_trace(buf.ctypes.data_as(C.c_void_p), ctas, slots) == 0
names = ["C20 work", "wait B1", "X1 work", "wait B2", "C1 work", "wait B3", "X2 work", "wait B4 (to next it)"]
for sub in (0, 1):
    st = buf[:, sub * 64: sub * 64 + 48].reshape(ctas, 6, 8)
    # deltas inside an iteration + to the next iteration's stamp 0
    d = np.diff(st, axis=2)                       # 7 deltas
    nxt = st[:, 1:, 0] - st[:, :-1, 7]            # wait B4
    print(f"slot {sub}: mean cycles over CTAs, iterations 2..5")
    for k in range(7):
        print(f"  {names[k]:24s} {d[:, 1:5, k].mean():8.0f}")
    print(f"  {names[7]:24s} {nxt[:, 1:4].mean():8.0f}")
    print(f"  iteration period         {(st[:, 1:, 0] - st[:, :-1, 0])[:, 1:4].mean():8.0f}")
