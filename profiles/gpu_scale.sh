#!/bin/bash
# Multi-GPU bench on one box: bash profiles/gpu_scale.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_smi_$N.txt 2>&1
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/scale_1_of_$N.json 2> gpurun_out/scale_1_of_$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "n=$n exit $?"
done
tail -c 1500 gpurun_out/scale_$N.json; tail -5 gpurun_out/scale_$N.err
