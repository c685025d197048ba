#!/bin/bash
# N-GPU measurement pass (N = $1): multi-GPU tests, all-reduce sweep, ping-pong, bench both arms at N.
set -u
N=${1:-2}
OUT=gpurun_out/multi$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== core test (fail fast)"
timeout 420 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -k symmetric_allreduce 2>&1 | tail -60 > $OUT/pytest_core.txt; tail -25 $OUT/pytest_core.txt
if ! grep -q "1 passed" $OUT/pytest_core.txt; then echo "CORE TEST FAILED -- skipping the rest"; ls -la $OUT; exit 1; fi
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -k "not symmetric_allreduce" 2>&1 | tail -60 > $OUT/pytest.txt; tail -25 $OUT/pytest.txt
echo "== sweep"; timeout 900 python bench/allreduce_sweep.py --gpus $N --max-mb ${2:-256} --out $OUT/sweep.json > $OUT/sweep.log 2>&1; tail -12 $OUT/sweep.log
if [ "$N" = "2" ]; then echo "== pingpong"; timeout 300 python bench/pingpong.py --out $OUT/pingpong.json > $OUT/pingpong.log 2>&1; tail -8 $OUT/pingpong.log; fi
for n in $(seq 1 $N); do
  if [ $n = 1 ] || [ $n = 2 ] || [ $n = 4 ] || [ $n = 8 ]; then
    echo "== bench ours N=$n"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 400 --warmup 20 > $OUT/bench_ours_$n.json 2> $OUT/bench_ours_$n.err; tail -c 1200 $OUT/bench_ours_$n.json; tail -3 $OUT/bench_ours_$n.err
    echo "== bench ref N=$n"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $n --steps 400 --warmup 20 > $OUT/bench_ref_$n.json 2> $OUT/bench_ref_$n.err; tail -c 1200 $OUT/bench_ref_$n.json; tail -3 $OUT/bench_ref_$n.err
  fi
done
echo "== resnet"; timeout 600 python bench/resnet_bench.py --gpus $N --out $OUT/resnet.json > $OUT/resnet.log 2>&1; tail -3 $OUT/resnet.log
NCCL_DEBUG=INFO timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 -m bench.nccl_probe > $OUT/nccl_info.log 2>&1; grep -iE "NVLS|nvls|Channel|Connected" $OUT/nccl_info.log | head -8
ls -la $OUT
