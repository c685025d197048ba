#!/bin/bash
# Round-end style check on 2 GPUs: full GPU test suite (as the driver runs it), smoke(), bench both arms N=1,2.
set -u
OUT=gpurun_out/final2
mkdir -p $OUT
echo "== pytest -m gpu (whole suite)"; timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > $OUT/pytest_gpu.txt; tail -6 $OUT/pytest_gpu.txt | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -2 $OUT/smoke.txt
for n in 1 2; do
for arm in ours reference; do
echo "== bench $arm N=$n"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --impl $arm --gpus $n --steps 400 --warmup 20 > $OUT/bench_${arm}_$n.json 2> $OUT/bench_${arm}_$n.err; tail -1 $OUT/bench_${arm}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['impl'], round(d['value']), round(d['ms_per_step']*1e3,1), 'us; e2e', round(d['e2e']['value']), d['clocks'])"; tail -2 $OUT/bench_${arm}_$n.err | cut -c1-200
done; done
echo "== default bench (no flags)"; timeout 300 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -1 $OUT/bench_default.json | cut -c1-300
echo "== ncu simt"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:convnet_step -s 6 -c 1 -o $OUT/prof_convnet_simt -f python bench.py --gpus 1 --steps 8 --warmup 3 --graph-chunk 1 --no-e2e > $OUT/ncu_simt.log 2>&1
ls $OUT
