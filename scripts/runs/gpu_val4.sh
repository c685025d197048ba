#!/bin/bash
set -u
OUT=gpurun_out/val4
mkdir -p $OUT
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > $OUT/pytest_single.txt; tail -12 $OUT/pytest_single.txt | cut -c1-200
echo "== tc diag"; timeout 300 python bench/tc_diag.py > $OUT/tc_diag.json 2>&1; python - <<'PY'
import json
t=open('gpurun_out/val4/tc_diag.json').read()
try:
    d=json.loads(t[t.index('{'):])
    for k in ('B1','B16','B128'):
        print(k, 'tc fwd', round(d[k]['tc']['fwd_max_abs_err'],5), {n:v for n,v in d[k]['tc']['grad_rel_err'].items()})
    print(d['train_curve_same_batch_lr0.01'])
except Exception as e: print('diag parse failed', e, t[-800:])
PY
for tc in 0 1; do
echo "== bench ours N=1 TC=$tc"; B200DIST_CONVNET_TC=$tc timeout 300 python bench.py --gpus 1 --steps 400 --warmup 20 > $OUT/bench_ours_1_tc$tc.json 2> $OUT/bench_ours_1_tc$tc.err; tail -1 $OUT/bench_ours_1_tc$tc.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"; tail -3 $OUT/bench_ours_1_tc$tc.err
done
echo "== kernel bench"; timeout 300 python bench/kernel_bench.py > $OUT/kernel_bench.json 2> $OUT/kernel_bench.err; python -c "
import json
d=json.load(open('$OUT/kernel_bench.json'))
for g in d['gemm']: print({k:(round(v,2) if isinstance(v,float) else v) for k,v in g.items() if k in ('M','N','K','us_median','cublas_us_median','tflops','frac_of_roofline')})
for c in d['convnet']: print(c['B'], round(c['fwd_bwd']['us_median'],1), round(c['fwd']['us_median'],1), round(c['sgd']['us_median'],1))
"
echo "== launches"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"convnet_step|allreduce_sgd" -s 10 -c 40 --csv --log-file $OUT/launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --graph-chunk 1 --no-e2e > $OUT/ncu_launch.log 2>&1; tail -4 $OUT/launches.csv | cut -c60-260
echo "== ncu TC"; B200DIST_CONVNET_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:convnet_step -s 6 -c 1 -o $OUT/prof_convnet_tc -f python bench.py --gpus 1 --steps 8 --warmup 3 --graph-chunk 1 --no-e2e > $OUT/ncu_tc.log 2>&1
echo "== ncu gemm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 1 -o $OUT/prof_gemm -f python bench/kernel_bench.py --gemm-only --big-only > $OUT/ncu_gemm.log 2>&1
echo "== sanitizer"; bash scripts/sanitize.sh 2>&1 | tail -12
ls -la $OUT | head -30
