"""Per-instruction stall samples from an ncu source-page CSV: top instructions and a phase profile."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]; A = ix["Address"]; SRC = ix["Source"]; EX = ix["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in body)
print("total samples", tot)
print("--- by stall reason ---")
agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stalls}
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:12]:
    print(f"{h:28s} {v:8d} {100*v/tot:5.1f}%")
print("--- top instructions ---")
top = sorted(body, key=lambda r: -int(r[S] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for r in top:
    reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{r[A][-5:]} {int(r[S]):6d} {100*int(r[S])/tot:4.1f}%  ex={r[EX]:>8s}  {r[SRC][:70]:70s} {reasons}")
if len(sys.argv) > 3:
    print("--- full listing ---")
    for r in body:
        reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{r[A][-5:]} {int(r[S] or 0):6d} ex={r[EX]:>8s} {r[SRC][:80]:80s} {reasons}")
