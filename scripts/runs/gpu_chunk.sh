#!/bin/bash
# One-GPU check of the chunk-graph executor + FlatSGD: numerics tests, then bench.py with chunk graphs on and off.
set -u
OUT=gpurun_out/chunk
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_optim.py -q -m gpu -p no:cacheprovider -k "executor or flat_sgd or train_loop" 2>&1 | tail -25 > $OUT/pytest.txt; tail -12 $OUT/pytest.txt
timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 50 > $OUT/bench_chunk4.json 2> $OUT/bench_chunk4.err; tail -c 1200 $OUT/bench_chunk4.json; tail -3 $OUT/bench_chunk4.err
B200DIST_EXEC_CHUNK=1 timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 50 > $OUT/bench_chunk1.json 2> $OUT/bench_chunk1.err; tail -c 1200 $OUT/bench_chunk1.json; tail -3 $OUT/bench_chunk1.err
B200DIST_EXEC_CHUNK=8 timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 50 --loader-buffers 16 > $OUT/bench_chunk8.json 2> $OUT/bench_chunk8.err; tail -c 1200 $OUT/bench_chunk8.json; tail -3 $OUT/bench_chunk8.err
