#!/bin/bash
# compute-sanitizer passes over small decodes that touch every kernel (profiles/sanitize_cases.py)
mkdir -p gpurun_out
rm -f gpurun_out/sanitizer_summary.txt
for tool in memcheck synccheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 3 python profiles/sanitize_cases.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitizer_summary.txt
  tail -4 gpurun_out/sanitizer_$tool.log
done
