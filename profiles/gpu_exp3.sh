#!/bin/bash
# Round-2 experiment 3: full ncu capture of k_imdct_fused (source-level stalls), fused variants with 14 / 15 warps.
mkdir -p gpurun_out
export NVB_BENCH_KERNELS_ONLY=1
B=nvorbis_b200/csrc/build
for spec in ${VARIANTS:-base:X=1 w15:NVB_LIB_PATH=$B/libnvb_w15.so w14:NVB_LIB_PATH=$B/libnvb_w14.so w12:NVB_LIB_PATH=$B/libnvb_w12.so}; do
  v=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/k_$v.json 2> gpurun_out/k_$v.err
  echo "$v $(cat gpurun_out/k_$v.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,1), "M f/s step", round(d["step_ms"]*1e3,1), "us single", round(d["step_ms_single_stream"]*1e3,1), "spec", round(d["k_spectrum_ms"]*1e3,1), "imdct", round(d["k_imdct_fused_ms"]*1e3,1), d["timing"]["repeats"], d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done
NVB_BENCH_MIN_S=0.001 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_imdct_fused -s 20 -c 1 -f -o gpurun_out/prof_fused python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_full_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
