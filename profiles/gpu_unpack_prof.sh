#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unpack.py -m gpu -x -q 2>&1 | tail -3
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,launch__registers_per_thread
timeout 300 ncu --metrics $M --clock-control none -k regex:k_unpack -s 4 -c 4 --csv --log-file gpurun_out/ncu_unpack.csv python profiles/prof_unpack.py 6 > gpurun_out/ncu_unpack.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ncu_unpack.csv')))
hdr=None
for r in rows:
    if len(r)>10 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d['ID'] in ('0','1'): print(d['ID'], d['Kernel Name'][:20], d['Metric Name'], d['Metric Value'])
PY
