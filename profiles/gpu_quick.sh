#!/bin/bash
# Quick GPU check during development: parity tests, kernel-resident timings (VARIANTS="name:ENV=val ..."), ncu counters of both kernels.
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
fi
export NVB_BENCH_KERNELS_ONLY=1
for spec in ${VARIANTS:-base:X=1}; do
  v=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/k_$v.json 2> gpurun_out/k_$v.err
  echo "$v $(cat gpurun_out/k_$v.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,1), "M f/s step", round(d["step_ms"]*1e3,1), "us single", round(d["step_ms_single_stream"]*1e3,1), "spec", round(d["k_spectrum_ms"]*1e3,1), "imdct", round(d["k_imdct_fused_ms"]*1e3,1), "one_kernel", round((d.get("one_kernel_ms") or 0)*1e3,1), d["timing"]["repeats"], d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done
if [ -z "$SKIP_NCU" ]; then
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
NVB_BENCH_MIN_S=0.001 timeout 300 ncu --metrics $M --clock-control none -k regex:k_ -s 12 -c 4 --csv --log-file gpurun_out/ncu_both.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_both.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ncu_both.csv')))
hdr=None; seen={}
for r in rows:
    if len(r)>10 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); k=d['Kernel Name'][:28]
        seen.setdefault((d['ID'],k),{})[d['Metric Name']]=d['Metric Value']
for (i,k),m in list(seen.items())[:4]:
    print(i,k,{a.split('__')[-1][:28]:b for a,b in m.items() if any(t in a for t in ('time_duration','inst_executed','cycles_active','issue_active','registers'))})
PY
fi
