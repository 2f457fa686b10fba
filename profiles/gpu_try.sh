#!/bin/bash
# Quick A/B of spectrum-kernel variants: parity tests, then the bench under each variant (VARIANTS="name:ENV=val ...").
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
fi
for spec in ${VARIANTS:-default:X=1}; do
  v=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$v.json")); print("$v", round(d["value"]/1e6,2), d["kernels"], round(d["roofline"]["frac"],3), round(d["e2e"]["value"]/1e6,2))
except Exception as e: print("$v failed", e)
PY
done
