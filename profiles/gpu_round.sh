#!/bin/bash
# One GPU-box session: parity tests, bench (ours + reference arm), ncu launch list, ncu full capture of both kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_imdct_fused -s 12 -c 2 -f -o gpurun_out/prof_fused python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spectrum -s 4 -c 1 -f -o gpurun_out/prof_spectrum python bench.py --steps 6 --warmup 3 >> gpurun_out/ncu_full.log 2>&1
timeout 600 python profiles/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
tail -5 gpurun_out/pytest.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err
