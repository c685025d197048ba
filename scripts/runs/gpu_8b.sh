#!/bin/bash
# 8-GPU scaling bench only (both arms at N=8,4,2,1 like the driver's scaling run) + multi-GPU tests at world 8.
set -u
OUT=gpurun_out/multi8b
mkdir -p $OUT
echo "== tests"; timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -k "fused_trainer or train_loop" 2>&1 | tail -20 > $OUT/pytest.txt; tail -4 $OUT/pytest.txt
for n in 8 4 2 1; do
  echo "== bench ours N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 400 --warmup 20 > $OUT/bench_ours_$n.json 2> $OUT/bench_ours_$n.err; tail -1 $OUT/bench_ours_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['impl'], round(d['value']), round(d['ms_per_step']*1e3,1), 'us; e2e', round(d['e2e']['value']), d['clocks'])"; tail -2 $OUT/bench_ours_$n.err | cut -c1-200
done
for n in 8; do
  echo "== bench ref N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $n --steps 400 --warmup 20 > $OUT/bench_ref_$n.json 2> $OUT/bench_ref_$n.err; tail -1 $OUT/bench_ref_$n.json | cut -c1-200
done
ls $OUT
