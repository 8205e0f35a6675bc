"""One chain call per config (FMA mode) for an ncu launch list: which kernel takes the time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rfanalyzer_b200 as rfa
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from bench_chain import CONFIGS
S = 1 << 26
stream = torch.cuda.Stream(); ctx = rfa.Context(0, stream)
which = [int(x) for x in os.environ.get("CONFIGS", "0,1,2").split(",")]
for idx in which:
    name, fmt, fs, mode, width, packet = CONFIGS[idx]
    off = fs // 10; mul = 256 if fmt == rfa.FMT_S16LE else 1
    comps = [(rfa.synth_step(off / fs), 60 * mul, rfa.synth_step(1000 / fs), 3_130_000 if mode in (2, 3) else 0),
             (rfa.synth_step((off + 1200) / fs), 20 * mul, 0, 0)]
    with torch.cuda.stream(stream):
        iq = torch.empty(S * rfa.BYTES_PER_SAMPLE[fmt], dtype=torch.uint8, device="cuda")
        rfa.synth_iq(ctx, fmt, S, iq, comps=comps, noise_shift=3)
        plan = rfa.ChainPlan(ctx, fmt, fs, 100_000_000, 100_000_000 + off, mode, width, packet, 1.0, rfa.SUM_FMA)
        audio = torch.empty(plan.max_audio(S), dtype=torch.float32, device="cuda")
        plan.process(iq, S, audio); plan.process(iq, S, audio)
        stream.synchronize()
print("done")
