#!/bin/bash
# Quick A/B of spectrum-kernel variants: parity tests, then the bench under each variant.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for v in default NT128 PLANES; do
  case $v in
    default) envs="";;
    NT128) envs="NVB_SPECTRUM_NT=128";;
    PLANES) envs="NVB_SPECTRUM_PLANES=1";;
  esac
  env $envs timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$v.json")); print("$v", d["value"], d["kernels"], d["roofline"]["frac"], d["e2e"]["value"])
except Exception as e: print("$v failed", e)
PY
done
