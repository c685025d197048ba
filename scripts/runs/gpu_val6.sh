#!/bin/bash
set -u
OUT=gpurun_out/val6
mkdir -p $OUT
echo "== cluster tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "cluster" 2>&1 | tail -60 > $OUT/pytest_cluster.txt; tail -25 $OUT/pytest_cluster.txt | cut -c1-220
echo "== kernel bench"; timeout 300 python bench/kernel_bench.py > $OUT/kernel_bench.json 2> $OUT/kernel_bench.err; tail -3 $OUT/kernel_bench.err; python -c "
import json
d=json.load(open('$OUT/kernel_bench.json'))
for c in d['convnet']: print(c['B'], 'C1', round(c['fwd_bwd']['us_median'],1), {k:round(v['us_median'],1) for k,v in c.items() if k.startswith('fwd_bwd_cluster')})
"
echo "== all single tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | cut -c1-200
