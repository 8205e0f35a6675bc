"""Instruction histogram of the largest loop of a kernel in a cuobjdump -sass listing."""
import re, sys, collections
path, pat = sys.argv[1], sys.argv[2]
fn = None; L = {}
for line in open(path):
    m = re.search(r'Function : (\S+)', line)
    if m: fn = m.group(1); L[fn] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
    if m and fn: L[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, lines in L.items():
    if pat not in fn: continue
    best = None
    for a, t in lines:
        m = re.search(r'BRA\S*\s+(?:!?U?P\d,?\s*)?(0x[0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            nfp = sum(1 for b, u in lines if lo <= b <= a and ('FADD2' in u or 'FADD ' in u))  # the FFT loop
            if best is None or nfp > best[2]: best = (lo, a, nfp)
    c = collections.Counter()
    for a, t in lines:
        if best[0] <= a <= best[1]:
            t = re.sub(r'^@!?U?P\d\s+', '', t); op = t.split()[0]
            if not op.startswith(('LDS', 'STS', 'LDG', 'STG', 'MUFU')): op = op.split('.')[0]
            c[op] += 1
    print(fn, 'total', len(lines), 'loop', hex(best[0]), hex(best[1]), sum(c.values()))
    print(sorted(c.items(), key=lambda x: -x[1]))
