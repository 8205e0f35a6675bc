import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import rfanalyzer_b200 as rfa
from oracle import oracle as O
s = torch.cuda.Stream(); ctx = rfa.Context(0, s)
n = 1 << 24
with torch.cuda.stream(s):
    x = ((torch.randint(1500, 2700, (n,), device="cuda", dtype=torch.int32) - 2048) * 16).to(torch.int16)
    cv = rfa.IqConverterInt16(ctx, O.synthetic_hb_kernel(47))
    for _ in range(3):
        cv.process(x.clone())
    ctx.sync()
print("ok")
