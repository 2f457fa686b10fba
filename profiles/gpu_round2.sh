#!/bin/bash
# Round-2 artefact run on one B200: parity tests, bench (ours + reference arm), ncu launch list of the bench command, ncu counter
# sets of the three hot kernels, environment probe.  Everything lands in gpurun_out/ and is copied into profiles/r02_* by hand.
mkdir -p gpurun_out
bash profiles/gpu_probe_env.sh > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
NVB_BENCH_KERNELS_ONLY=1 NVB_BENCH_MIN_S=0.001 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
NVB_BENCH_KERNELS_ONLY=1 NVB_BENCH_MIN_S=0.001 timeout 300 ncu --metrics $M --clock-control none -k regex:k_ -s 12 -c 2 --csv --log-file gpurun_out/ncu_metrics.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_metrics.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:k_unpack -s 2 -c 1 --csv --log-file gpurun_out/ncu_metrics_unpack.csv python profiles/prof_unpack.py 4 >> gpurun_out/ncu_metrics.log 2>&1
python - <<'PY'
import csv, json
out = []
for fn in ("gpurun_out/ncu_metrics.csv", "gpurun_out/ncu_metrics_unpack.csv"):
    rows = list(csv.reader(open(fn))); hdr = None; seen = {}
    for r in rows:
        if len(r) > 10 and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r)); seen.setdefault((d["ID"], d["Kernel Name"]), {})[d["Metric Name"]] = (d["Metric Unit"], d["Metric Value"])
    for (i, k), m in seen.items():
        out.append(f"-- {k}")
        out += [f"   {a} {u} {v}" for a, (u, v) in sorted(m.items())]
open("gpurun_out/ncu_raw_metrics.txt", "w").write("\n".join(out) + "\n")
d = json.load(open("gpurun_out/bench.json"))
print("value", round(d["value"] / 1e6, 1), "spec", d["kernels"]["k_spectrum_ms"], "imdct", d["kernels"]["k_imdct_fused_ms"], "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"] / 1e6, 2),
      "s16", round(d["e2e"]["s16_value"] / 1e6, 2), "strong", round(d["strong_64k"]["value"] / 1e6, 1), round(d["strong_64k"]["roofline_k_imdct_fused"]["frac"], 3), "ogg", {k: round(v["frames_per_s"] / 1e6, 2) for k, v in d["ogg_to_pcm"].items() if isinstance(v, dict)},
      "cfg", {k[:10]: (round(v["frames_per_s"] / 1e6, 1), round(v["imdct_roofline_frac"], 3)) for k, v in d["other_configs"].items()})
PY
