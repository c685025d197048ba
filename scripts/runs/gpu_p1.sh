#!/bin/bash
# Cheap one-GPU check after a ConvNet kernel change: numerics tests + warm step matrix.
set -u
OUT=gpurun_out/p1
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "convnet or fused or train or executor or cluster" 2>&1 | tail -15 > $OUT/pytest.txt; tail -8 $OUT/pytest.txt
timeout 300 python bench/step_bench.py > $OUT/step_bench.json 2> $OUT/step_bench.err; tail -c 2500 $OUT/step_bench.json; tail -3 $OUT/step_bench.err
