#!/bin/bash
# round 2, call 1: validate the re-worked bench timing window at the driver's step count (20) and at 400 steps, N=1 and N=2
set -x
mkdir -p gpurun_out/r2_c1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_c1/smi.txt
for K in 20 400; do
  timeout 300 python bench.py --gpus 1 --steps $K --warmup 5 > gpurun_out/r2_c1/n1_k$K.json 2> gpurun_out/r2_c1/n1_k$K.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps $K --warmup 5 > gpurun_out/r2_c1/n2_k$K.json 2> gpurun_out/r2_c1/n2_k$K.err
done
tail -n 3 gpurun_out/r2_c1/*.json
tail -n 5 gpurun_out/r2_c1/*.err
