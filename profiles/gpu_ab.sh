#!/bin/bash
# A/B of library build variants (VARIANTS="name:ENV=val ..."): headline kernel timings + the mixed-window config.
mkdir -p gpurun_out
export NVB_BENCH_KERNELS_ONLY=1
for spec in ${VARIANTS:-base:X=1}; do
  v=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/k_$v.json 2> gpurun_out/k_$v.err
  echo "$v $(cat gpurun_out/k_$v.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,1), "M f/s step", round(d["step_ms"]*1e3,1), "spec", round(d["k_spectrum_ms"]*1e3,1), "imdct", round(d["k_imdct_fused_ms"]*1e3,1))' 2>&1 | tail -1) | cfg3: $(env $envs python profiles/prof_config3.py 2>&1 | tail -1)"
done
