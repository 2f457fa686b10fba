#!/bin/bash
# ncu full capture of the spectrum kernel (source-level) + a short bench; used to find where its instructions go.
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spectrum -s 4 -c 1 -o gpurun_out/prof_spectrum python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_imdct_fused -s 12 -c 1 -o gpurun_out/prof_fused python bench.py --steps 6 --warmup 3 >> gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; ls -la gpurun_out
