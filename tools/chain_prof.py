"""ncu target: one warmed-up call of the IQ->audio chain per config (C2 wFM, C4 nFM / USB), 2^24 samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rfanalyzer_b200 as rfa
S = 1 << int(os.environ.get("LOG2", "24"))
which = os.environ.get("CFG", "all")
CFG = {"wfm": (rfa.FMT_U8, 2_400_000, rfa.MODE_WFM, 100_000, 8192), "nfm": (rfa.FMT_S16LE, 10_000_000, rfa.MODE_NFM, 10_000, 65536),
       "usb": (rfa.FMT_S16LE, 10_000_000, rfa.MODE_USB, 2_800, 65536)}
stream = torch.cuda.Stream(); ctx = rfa.Context(0, stream)
for name, (fmt, fs, mode, width, packet) in CFG.items():
    if which not in ("all", name): continue
    off = fs // 10; mul = 256 if fmt == 2 else 1
    comps = [(rfa.synth_step(off / fs), 60 * mul, rfa.synth_step(1000 / fs), 3_130_000 if mode in (2, 3) else 0), (rfa.synth_step((off + 1200) / fs), 20 * mul, 0, 0)]
    with torch.cuda.stream(stream):
        iq = torch.empty(S * rfa.BYTES_PER_SAMPLE[fmt], dtype=torch.uint8, device="cuda")
        rfa.synth_iq(ctx, fmt, S, iq, comps=comps, noise_shift=3)
        plan = rfa.ChainPlan(ctx, fmt, fs, 100_000_000, 100_000_000 + off, mode, width, packet, 1.0, rfa.SUM_FMA)
        audio = torch.empty(plan.max_audio(S), dtype=torch.float32, device="cuda")
        plan.process(iq, S, audio); stream.synchronize()
        torch.cuda.profiler.start()
        plan.process(iq, S, audio); stream.synchronize()
        torch.cuda.profiler.stop()
print("done")
