#!/bin/bash
# ncu full capture of the spectrum kernel (source-level); env passes through (e.g. NVB_SPECTRUM_NT=128)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spectrum -s 4 -c 1 -o gpurun_out/prof_spectrum -f python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
