#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: stall reasons, instruction mix, hot instructions.
A dump can hold several functions (template instantiations of one kernel): only the FIRST function section is summarised --
round 1 summed them all, which multiplied "total warp inst" by the number of instantiations and listed hot rows twice."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
def num(x):
    try: return int(float(x))
    except Exception: return 0
tot = Counter(); total = 0; data = []; mix = Counter()
seen = set()
for r in rows[2:]:
    if len(r) >= 1 and r[0] in ('Function Name', 'Address') and data: break          # the next function's section starts
    if len(r) < len(hdr) or r[0] == 'Address' or not r[0].startswith('0x'): continue
    if r[0] in seen: continue                                                         # an address listed twice
    seen.add(r[0])
    s = num(r[ix['# Samples']]); total += s
    for k in stalls: tot[k] += num(r[ix[k]])
    n = num(r[ix['Instructions Executed']])
    src = r[ix['Source']].strip()
    parts = src.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    mix[op.split('.')[0]] += n
    data.append((s, src, n, r))
print('total samples', total)
for k, v in tot.most_common(10): print(f'  {k:28s} {v:8d} {100*v/max(total,1):5.1f}%')
tn = sum(mix.values()); print('total warp inst', tn)
for k, v in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22): print(f'  {k:12s} {v:9d} {100*v/tn:5.1f}%')
print('hot instructions (samples, executed, sass, top stalls)')
for s, src, n, r in sorted(data, key=lambda x: -x[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    top = sorted(((num(r[ix[k]]), k) for k in stalls), reverse=True)[:2]
    extra = ''
    if 'L1 Wavefronts Shared Excessive' in ix and num(r[ix['L1 Wavefronts Shared Excessive']]): extra = f" smem_excess={r[ix['L1 Wavefronts Shared Excessive']]}/{r[ix['L1 Wavefronts Shared']]}"
    print(f'  {s:6d} {n:8d} {src[:64]:64s} {top}{extra}')
