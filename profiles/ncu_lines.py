#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA source line: instructions, stall samples."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
agg = defaultdict(lambda: [0, 0]); src = {}; cur = None; fpath = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] == 'Line No': continue
    if r[0] != '': cur = (fpath, int(r[0])); src[cur] = r[1]
    else:
        try: n = int(r[7]); s = int(r[6])
        except Exception: continue
        agg[cur][0] += n; agg[cur][1] += s
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('total warp instructions', tot, 'samples', ts)
for k in sorted(agg, key=lambda k: (k[0], k[1])):
    n, s = agg[k]
    if n > thr or s > ts / 100: print(f'{n:9d} {100*n/tot:5.1f}% {s:5d} {100*s/ts:5.1f}% {k[0][:18]:18s}:{k[1]:4d} {src[k][:100]}')
