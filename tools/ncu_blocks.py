"""Where a kernel's executed instructions and stall samples sit: contiguous SASS blocks with (nearly) the same execution
count, from an `ncu -i rep --page source --csv --kernel-name regex:NAME` listing (first kernel instance in the file)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
EX, SRC, S, A = ix['Instructions Executed'], ix['Source'], ix['# Samples'], ix['Address']
body = []
for r in rows[2:]:
    if r and r[0] in ('Kernel Name', 'Address'):
        if r[0] == 'Kernel Name': break
        continue
    if len(r) > EX: body.append(r)
tot = sum(int(r[EX] or 0) for r in body); ts = sum(int(r[S] or 0) for r in body)
print('total warp instructions', tot, 'stall samples', ts, 'SASS lines', len(body))
blocks = []; cur = None
for r in body:
    e = int(r[EX] or 0); op = r[SRC].split()[0] if r[SRC] else ''
    if op.startswith('@'): op = (r[SRC].split() + [''])[1]
    if cur and abs(e - cur[0]) <= 0.03 * max(e, cur[0], 1):
        cur[1] += 1; cur[2] += e; cur[3] += int(r[S] or 0); cur[5] = r[A][-5:]; cur[6].append(op)
    else:
        cur = [e, 1, e, int(r[S] or 0), r[A][-5:], r[A][-5:], [op]]; blocks.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.015
for b in blocks:
    if b[2] > tot * thr:
        c = collections.Counter(x.split('.')[0] for x in b[6])
        print(f"{b[4]}-{b[5]} n={b[1]:4d} exec/inst={b[0]:8d} inst={100*b[2]/tot:5.1f}% samples={100*b[3]/ts:5.1f}%", c.most_common(7))
