#!/bin/bash
# Multi-GPU bench on one box: bash profiles/gpu_scale.sh "1 2 8"   (run under gpurun --gpus N, N = the largest)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_smi.txt 2>&1
for n in ${1:-1 2}; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "n=$n exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1]); print($n, round(d["value"]/1e6,2), "M frames/s device", round(d["e2e"]["value"]/1e6,2), "e2e", d["kernels"])
except Exception as e: print("failed", e)
PY
done
