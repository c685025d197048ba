#!/bin/bash
mkdir -p gpurun_out/r2_c5
python scripts/bt_stage_probe.py 64 > gpurun_out/r2_c5/stages.txt 2>&1
cat gpurun_out/r2_c5/stages.txt
timeout 900 python -m pytest tests/test_gpu_batched.py -q --timeout 600 -x > gpurun_out/r2_c5/batched.txt 2>&1; echo "batched rc=$?" >> gpurun_out/r2_c5/batched.txt
tail -n 25 gpurun_out/r2_c5/batched.txt
cp gpurun_out/batched_diag_*.json gpurun_out/r2_c5/ 2>/dev/null
timeout 600 python bench/batched_bench.py --batch 1024 4096 --steps 20 --out gpurun_out/r2_c5/batched_bench.json > gpurun_out/r2_c5/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/r2_c5/bench.txt
tail -n 12 gpurun_out/r2_c5/bench.txt
for K in 20 400; do timeout 300 python bench.py --gpus 1 --steps $K --warmup 5 > gpurun_out/r2_c5/n1_k$K.json 2> gpurun_out/r2_c5/n1_k$K.err; tail -c 900 gpurun_out/r2_c5/n1_k$K.json; tail -n 3 gpurun_out/r2_c5/n1_k$K.err; done
