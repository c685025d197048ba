#!/bin/bash
set -u
OUT=gpurun_out/val2
mkdir -p $OUT
echo "== single-GPU new tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "native_executor or train_loop or fused_trainer" 2>&1 | tail -30 > $OUT/pytest_single.txt; tail -6 $OUT/pytest_single.txt
echo "== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > $OUT/pytest_multi.txt; tail -8 $OUT/pytest_multi.txt
echo "== sweep"; timeout 300 python bench/allreduce_sweep.py --gpus 2 --max-mb 64 --out $OUT/sweep.json > $OUT/sweep.log 2>&1; tail -4 $OUT/sweep.log | cut -c1-400
for n in 1 2; do
echo "== bench ours N=$n"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 400 --warmup 20 > $OUT/bench_ours_$n.json 2> $OUT/bench_ours_$n.err; tail -1 $OUT/bench_ours_$n.json | cut -c1-300; tail -3 $OUT/bench_ours_$n.err
done
echo "== resnet"; timeout 400 python bench/resnet_bench.py --gpus 2 --out $OUT/resnet.json > $OUT/resnet.log 2>&1; tail -2 $OUT/resnet.log | cut -c1-400
