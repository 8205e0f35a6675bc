"""Phase timeline of the single-frame spectrum kernel (-DRFA_TRACE build; DUAL=1 in the environment of this script selects the lab build's dual-frame kernel): cycles per phase."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DUAL = os.environ.get("DUAL", "0")
import numpy as np, torch
import rfanalyzer_b200 as rfa
from rfanalyzer_b200 import _lib
from oracle import oracle as O

N = int(os.environ.get("N", "4096")); F = (1 << 24) // N
stream = torch.cuda.Stream(); ctx = rfa.Context(0, stream)
if DUAL != "0":
    ctx.set_option("kernel", 1)
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
ctas, slots = (296 if DUAL == "0" else 148), 128
buf = np.zeros((ctas, slots), np.int64)
assert lib.rfa_debug_trace(buf.ctypes.data_as(C.c_void_p), ctas, slots) == 0
t0 = buf[:, 0].min()
rel = buf - t0
print("kernel start spread (cycles): max", rel[:, 0].max())
print("prologue: mean %.0f" % (buf[:, 1] - buf[:, 0]).mean())
print("loop total: mean %.0f  min %.0f max %.0f" % ((buf[:, 2] - buf[:, 1]).mean(), (buf[:, 2] - buf[:, 1]).min(), (buf[:, 2] - buf[:, 1]).max()))
print("peak atomics: mean %.0f" % (buf[:, 3] - buf[:, 2]).mean())
print("avg+end: mean %.0f max %.0f" % ((buf[:, 4] - buf[:, 3]).mean(), (buf[:, 4] - buf[:, 3]).max()))
print("kernel end (rel to first start): mean %.0f max %.0f" % (rel[:, 4].mean(), rel[:, 4].max()))
names = ["first pass (convert+bfly A)+scatter1", "barrier1 wait", "pass B (gather+tw+bfly)+scatter2", "barrier2 wait",
         "(unused)", "pass C gather+bfly", "emit"]
its = 13
ph = np.zeros((its, 7))
for it in range(its):
    b = 8 + 8 * it
    st = buf[:, b:b + 8]
    d = np.diff(st[:, [0, 1, 2, 3, 4, 4, 6, 7]], axis=1)
    ph[it] = d.mean(axis=0)
print("per-iteration phase means over CTAs (cycles), iterations 0..%d" % (its - 1))
for k, n in enumerate(names):
    print(f"  {n:40s} mean {ph[2:, k].mean():8.0f}   (it0 {ph[0, k]:6.0f}, it1 {ph[1, k]:6.0f})")
print("  iteration total mean %.0f" % np.diff(buf[:, 8:8 + 8 * its:8], axis=1).mean())
# co-residency: which CTAs share an SM, and their phase offset
sm = buf[:, 5]
pairs = {}
for c in range(ctas): pairs.setdefault(int(sm[c]), []).append(c)
offs = [abs(int(buf[v[0], 8 + 8 * 6] - buf[v[1], 8 + 8 * 6])) for v in pairs.values() if len(v) == 2]
print("SMs with 2 CTAs:", len(offs), " |iteration-6 start offset| between co-resident CTAs: mean %.0f median %.0f" % (np.mean(offs), np.median(offs)))
print("--- co-resident CTA pairs: start offsets and ends (cycles, same SM clock) ---")
k = 0
for s_, v in sorted(pairs.items()):
    if len(v) == 2 and k < 10:
        a, b = v
        base = min(buf[a, 0], buf[b, 0])
        print(f"SM {s_:3d} CTAs {a:3d},{b:3d}: start {buf[a,0]-base:6d} {buf[b,0]-base:6d}  loop start {buf[a,1]-base:6d} {buf[b,1]-base:6d}  loop end {buf[a,2]-base:6d} {buf[b,2]-base:6d}  end {buf[a,4]-base:6d} {buf[b,4]-base:6d}")
        k += 1
starts = [abs(int(buf[v[0], 0] - buf[v[1], 0])) for v in pairs.values() if len(v) == 2]
print("start offset between co-resident CTAs: mean %.0f min %d max %d" % (np.mean(starts), min(starts), max(starts)))
ends = [max(int(buf[v[0], 4]), int(buf[v[1], 4])) - min(int(buf[v[0], 0]), int(buf[v[1], 0])) for v in pairs.values() if len(v) == 2]
print("per-SM busy span: mean %.0f min %d max %d" % (np.mean(ends), min(ends), max(ends)))
it_tot = np.diff(buf[:, 8:8 + 8 * its:8], axis=1)
print("iteration time by index (mean over CTAs):", np.round(it_tot.mean(axis=0)).astype(int))
