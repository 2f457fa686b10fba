#!/bin/bash
# compute-sanitizer passes over a small decode (smoke: 1test full + 60 frames of 3test, exact and fused paths)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitizer_summary.txt
  tail -4 gpurun_out/sanitizer_$tool.log
done
