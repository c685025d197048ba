#!/bin/bash
set -u
OUT=gpurun_out/val5
mkdir -p $OUT
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > $OUT/pytest_single.txt; tail -6 $OUT/pytest_single.txt | cut -c1-200
for pdl in 1 0; do
echo "== bench ours N=1 PDL=$pdl"; B200DIST_PDL=$pdl timeout 300 python bench.py --gpus 1 --steps 400 --warmup 20 > $OUT/bench_ours_1_pdl$pdl.json 2> $OUT/bench_ours_1_pdl$pdl.err; tail -1 $OUT/bench_ours_1_pdl$pdl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"; tail -3 $OUT/bench_ours_1_pdl$pdl.err
done
echo "== bench ref"; timeout 300 python bench.py --impl reference --gpus 1 --steps 400 --warmup 20 > $OUT/bench_ref_1.json 2> $OUT/bench_ref_1.err; tail -1 $OUT/bench_ref_1.json | cut -c1-200
echo "== kernel bench"; timeout 300 python bench/kernel_bench.py > $OUT/kernel_bench.json 2> $OUT/kernel_bench.err; python -c "
import json
d=json.load(open('$OUT/kernel_bench.json'))
for g in d['gemm']: print({k:(round(v,2) if isinstance(v,float) else v) for k,v in g.items() if k in ('M','N','K','us_median','cublas_us_median','tflops','frac_of_roofline')})
for c in d['convnet']: print(c['B'], round(c['fwd_bwd']['us_median'],1), round(c['fwd']['us_median'],1), round(c['sgd']['us_median'],1))
"
echo "== ncu simt"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:convnet_step -s 6 -c 1 -o $OUT/prof_convnet_simt -f python bench.py --gpus 1 --steps 8 --warmup 3 --graph-chunk 1 --no-e2e > $OUT/ncu_simt.log 2>&1
ls $OUT
