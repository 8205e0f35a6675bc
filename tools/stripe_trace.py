"""Phase totals of the stripe resampler (-DRFA_STRIPE_TRACE build, RFA_B200_LIB=.../librfa_b200_trace.so)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import rfanalyzer_b200 as rfa
from rfanalyzer_b200 import _lib
S = 1 << 24
fmt, fs, mode, width, packet = rfa.FMT_S16LE, 10_000_000, rfa.MODE_NFM, 10_000, 65536
stream = torch.cuda.Stream(); ctx = rfa.Context(0, stream)
off = fs // 10
with torch.cuda.stream(stream):
    iq = torch.empty(S * 4, dtype=torch.uint8, device="cuda")
    rfa.synth_iq(ctx, fmt, S, iq)
    plan = rfa.ChainPlan(ctx, fmt, fs, 100_000_000, 100_000_000 + off, mode, width, packet, 1.0, rfa.SUM_FMA)
    audio = torch.empty(plan.max_audio(S), dtype=torch.float32, device="cuda")
    plan.process(iq, S, audio); stream.synchronize()
    lib = C.CDLL(_lib.LIB_PATH)
    buf = (C.c_ulonglong * 8)()
    lib.rfa_debug_stripe_trace(buf)
    plan.process(iq, S, audio); stream.synchronize()
    lib.rfa_debug_stripe_trace(buf)
names = ["prologue", "tile top (barrier, slots)", "wait bulk copy", "decode", "barrier after decode", "dot products"]
tot = sum(buf[:6])
for n, v in zip(names, buf[:6]):
    print("%-28s %10.0f cycles per CTA  %5.1f%%" % (n, v / 148, 100.0 * v / tot))
print("sum per CTA: %.0f cycles" % (tot / 148))
