#!/bin/bash
# Round-2 experiment 1: k_spectrum_wf (1 / 2 / 4 warps per frame) against k_spectrum_run: parity, kernel-resident timings, ncu counters.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
export NVB_BENCH_KERNELS_ONLY=1
for spec in ${VARIANTS:-wf2:NVB_WF_WPF=2 wf1:NVB_WF_WPF=1 wf4:NVB_WF_WPF=4 run:NVB_SPECTRUM_RUN=1}; do
  v=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/k_$v.json 2> gpurun_out/k_$v.err
  echo "$v $(cat gpurun_out/k_$v.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,1), "M f/s step", round(d["step_ms"]*1e3,1), "us single", round(d["step_ms_single_stream"]*1e3,1), "spec", round(d["k_spectrum_ms"]*1e3,1), "imdct", round(d["k_imdct_fused_ms"]*1e3,1), d["timing"]["repeats"], d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__occupancy_limit_warps,launch__occupancy_limit_blocks,launch__occupancy_limit_barriers,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
for spec in wf2:NVB_WF_WPF=2 wf1:NVB_WF_WPF=1; do
  v=${spec%%:*}; envs=${spec#*:}
  env $envs NVB_BENCH_MIN_S=0.001 timeout 300 ncu --metrics $M --clock-control none -k regex:k_spectrum -s 8 -c 2 --csv --log-file gpurun_out/ncu_$v.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_$v.log 2>&1
done
NVB_BENCH_MIN_S=0.001 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spectrum -s 8 -c 1 -f -o gpurun_out/prof_wf2 python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
