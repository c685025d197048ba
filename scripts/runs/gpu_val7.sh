#!/bin/bash
set -u
OUT=gpurun_out/val7
mkdir -p $OUT
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider 2>&1 | tail -12 > $OUT/pytest_single.txt; tail -5 $OUT/pytest_single.txt | cut -c1-200
echo "== bench"; timeout 300 python bench.py --gpus 1 --steps 400 --warmup 20 > $OUT/bench_ours_1.json 2> $OUT/bench_ours_1.err; tail -1 $OUT/bench_ours_1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1e3,2), 'us; e2e', round(d['e2e']['value']))"; tail -2 $OUT/bench_ours_1.err
echo "== kernel bench"; timeout 300 python bench/kernel_bench.py > $OUT/kernel_bench.json 2> $OUT/kernel_bench.err; tail -2 $OUT/kernel_bench.err; python -c "
import json
d=json.load(open('$OUT/kernel_bench.json'))
for c in d['convnet']: print(c['B'], 'C1', round(c['fwd_bwd']['us_median'],1), {k:round(v['us_median'],1) for k,v in c.items() if k.startswith('fwd_bwd_cluster')})
"
echo "== ncu cluster4"; B200DIST_CONVNET_CLUSTER=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:convnet_cluster -s 4 -c 1 -o $OUT/prof_cluster4 -f python -c "
import torch, sys
sys.path.insert(0,'.')
from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer
tr = FusedTrainer(32, device='cuda:0', cluster=4, use_graph=False)
x = torch.randn(32,1,28,28,device='cuda'); y = torch.randint(0,10,(32,),device='cuda')
for _ in range(8): tr.step(x, y)
torch.cuda.synchronize()
" > $OUT/ncu_cluster.log 2>&1; tail -2 $OUT/ncu_cluster.log
ls $OUT
