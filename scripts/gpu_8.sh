#!/bin/bash
# 8-GPU pass (expensive: charged 8x) -- core tests, sweep to 1 GiB, scaling bench N=4,8 (both arms), ResNet-18.
set -u
N=8
OUT=gpurun_out/multi8
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/nproc.txt
echo "== core test (fail fast)"
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -k "symmetric_allreduce" 2>&1 | tail -40 > $OUT/pytest_core.txt; tail -12 $OUT/pytest_core.txt
if ! grep -q "1 passed" $OUT/pytest_core.txt; then echo "CORE TEST FAILED -- skipping the rest"; ls -la $OUT; exit 1; fi
echo "== trainer + ddp tests"; timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -k "fused_trainer or average_gradients" 2>&1 | tail -30 > $OUT/pytest.txt; tail -8 $OUT/pytest.txt
echo "== sweep"; timeout 420 python bench/allreduce_sweep.py --gpus 8 --max-mb 1024 --out $OUT/sweep.json > $OUT/sweep.log 2>&1; tail -6 $OUT/sweep.log | cut -c1-300
for n in 8 4; do
  echo "== bench ours N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 400 --warmup 20 > $OUT/bench_ours_$n.json 2> $OUT/bench_ours_$n.err; tail -1 $OUT/bench_ours_$n.json | cut -c1-600; tail -2 $OUT/bench_ours_$n.err
  echo "== bench ref N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $n --steps 400 --warmup 20 > $OUT/bench_ref_$n.json 2> $OUT/bench_ref_$n.err; tail -1 $OUT/bench_ref_$n.json | cut -c1-600; tail -2 $OUT/bench_ref_$n.err
done
echo "== resnet"; timeout 400 python bench/resnet_bench.py --gpus 8 --out $OUT/resnet.json > $OUT/resnet.log 2>&1; tail -2 $OUT/resnet.log | cut -c1-400
ls -la $OUT
