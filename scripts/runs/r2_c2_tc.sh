#!/bin/bash
# round 2, call 2: tensor-core layout probes + batched engine numerics + first timings (1 GPU)
set -x
mkdir -p gpurun_out/r2_c2
timeout 600 python -m pytest tests/test_gpu_tc_probe.py -q -x --timeout 300 > gpurun_out/r2_c2/probe.txt 2>&1; echo "probe rc=$?" >> gpurun_out/r2_c2/probe.txt
timeout 900 python -m pytest tests/test_gpu_batched.py -q --timeout 600 > gpurun_out/r2_c2/batched.txt 2>&1; echo "batched rc=$?" >> gpurun_out/r2_c2/batched.txt
timeout 600 python bench/batched_bench.py --batch 1024 4096 --steps 20 --out gpurun_out/r2_c2/batched_bench.json > gpurun_out/r2_c2/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/r2_c2/bench.txt
tail -n 30 gpurun_out/r2_c2/probe.txt gpurun_out/r2_c2/batched.txt gpurun_out/r2_c2/bench.txt
