#!/bin/bash
set -u
OUT=gpurun_out/val3
mkdir -p $OUT
echo "== single-GPU tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > $OUT/pytest_single.txt; tail -12 $OUT/pytest_single.txt
echo "== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > $OUT/pytest_multi.txt; tail -8 $OUT/pytest_multi.txt
for tc in 0 1; do for n in 1 2; do
echo "== bench ours N=$n TC=$tc"; B200DIST_CONVNET_TC=$tc timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 400 --warmup 20 > $OUT/bench_ours_${n}_tc$tc.json 2> $OUT/bench_ours_${n}_tc$tc.err; tail -1 $OUT/bench_ours_${n}_tc$tc.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"; tail -3 $OUT/bench_ours_${n}_tc$tc.err
done; done
echo "== ncu TC kernel"
B200DIST_CONVNET_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:convnet_step -s 6 -c 2 -o $OUT/prof_convnet_tc -f python bench.py --gpus 1 --steps 8 --warmup 3 --graph-chunk 1 --no-e2e > $OUT/ncu_tc.log 2>&1
B200DIST_CONVNET_TC=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:convnet_step -s 6 -c 2 -o $OUT/prof_convnet_simt -f python bench.py --gpus 1 --steps 8 --warmup 3 --graph-chunk 1 --no-e2e > $OUT/ncu_simt.log 2>&1
timeout 300 python bench/kernel_bench.py > $OUT/kernel_bench.json 2> $OUT/kernel_bench.err
ls -la $OUT
