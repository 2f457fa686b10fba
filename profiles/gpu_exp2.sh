#!/bin/bash
# Round-2 experiment 2: k_spectrum_wf v2 (runs of 16, walked floor, branch-free gather) + fused-kernel build variants (warps / no register prefetch).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
export NVB_BENCH_KERNELS_ONLY=1
B=nvorbis_b200/csrc/build
for spec in ${VARIANTS:-wf1:NVB_WF_WPF=1 wf2:NVB_WF_WPF=2 w16np:NVB_LIB_PATH=$B/libnvb_w16np.so w20np:NVB_LIB_PATH=$B/libnvb_w20np.so w24np:NVB_LIB_PATH=$B/libnvb_w24np.so}; do
  v=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/k_$v.json 2> gpurun_out/k_$v.err
  echo "$v $(cat gpurun_out/k_$v.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,1), "M f/s step", round(d["step_ms"]*1e3,1), "us single", round(d["step_ms_single_stream"]*1e3,1), "spec", round(d["k_spectrum_ms"]*1e3,1), "imdct", round(d["k_imdct_fused_ms"]*1e3,1), d["timing"]["repeats"], d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
NVB_BENCH_MIN_S=0.001 timeout 300 ncu --metrics $M --clock-control none -k regex:k_spectrum -s 8 -c 2 --csv --log-file gpurun_out/ncu_wf1.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_wf1.log 2>&1
NVB_BENCH_MIN_S=0.001 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spectrum -s 8 -c 1 -f -o gpurun_out/prof_wf1 python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
