#!/bin/bash
# 8 GPUs, ordered by importance; every part independent and bounded.
O=gpurun_out/r2_c11; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    lb=d['config'].get('large_batch') or {}
    print(sys.argv[1], "N=%d value %.3fM us/step %.1f e2e %s large_batch %s"%(d['n_gpus'], d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value'] and "%.3fM"%(d['e2e']['value']/1e6), lb.get('samples_per_s') and "%.1fM"%(lb['samples_per_s']/1e6)))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
bench() { name=$1; n=$2; port=$3; shift 3; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; summ $O/$name.json; }
# ---- 1. the driver's curve: N = 8 alone, then 4 | 2 | 1 side by side on disjoint GPUs
bench n8_k20 8 29701 --steps 20 --warmup 5
bench n8_k400 8 29702 --steps 400 --warmup 5 --large-batch 0 --no-e2e
B200DIST_WIRE=bf16 bench n8_k400_bf16wire 8 29703 --steps 400 --warmup 5 --large-batch 0 --no-e2e
(CUDA_VISIBLE_DEVICES=0,1,2,3 bench n4_k20 4 29704 --steps 20 --warmup 5) &
(CUDA_VISIBLE_DEVICES=4,5 bench n2_k20 2 29705 --steps 20 --warmup 5) &
(CUDA_VISIBLE_DEVICES=6 timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/n1_k20.json 2> $O/n1_k20.err; summ $O/n1_k20.json) &
wait
# ---- 2. tests: everything at 8 in one launch, then the odd worlds (7 | then 3 and 5 side by side | then 6)
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 800 -s -k "world8_suite" > $O/suite8.txt 2>&1; echo "suite8 rc=$?" | tee -a $O/suite8.txt; grep -E "SUITE|passed|failed|Error" $O/suite8.txt | tail -16
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -k "odd_worlds and 7" > $O/odd7.txt 2>&1; echo "odd7 rc=$?" | tee -a $O/odd7.txt; tail -n 3 $O/odd7.txt
(CUDA_VISIBLE_DEVICES=0,1,2 timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -k "odd_worlds and 3" > $O/odd3.txt 2>&1; echo "odd3 rc=$?" >> $O/odd3.txt) &
(CUDA_VISIBLE_DEVICES=3,4,5,6,7 timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -k "odd_worlds and 5" > $O/odd5.txt 2>&1; echo "odd5 rc=$?" >> $O/odd5.txt) &
wait
tail -n 3 $O/odd3.txt $O/odd5.txt
(CUDA_VISIBLE_DEVICES=0,1,2,3,4,5 timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -k "odd_worlds and 6" > $O/odd6.txt 2>&1; echo "odd6 rc=$?" >> $O/odd6.txt) &
(CUDA_VISIBLE_DEVICES=6,7 timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -k "subgroup" > $O/subgroup.txt 2>&1; echo "subgroup rc=$?" >> $O/subgroup.txt) &
wait
tail -n 3 $O/odd6.txt $O/subgroup.txt
# ---- 3. all-reduce sweeps -> variant tables (8; then 4 and 3 side by side)
timeout 900 python bench/allreduce_sweep.py --gpus 8 --max-mb 1024 --emit-table --out $O/sweep_8.json > $O/sweep_8.txt 2>&1; tail -n 2 $O/sweep_8.txt | cut -c1-300
(CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 python bench/allreduce_sweep.py --gpus 4 --max-mb 64 --emit-table --out $O/sweep_4.json > $O/sweep_4.txt 2>&1) &
sleep 2
(CUDA_VISIBLE_DEVICES=4,5,6 timeout 600 python bench/allreduce_sweep.py --gpus 3 --max-mb 64 --out $O/sweep_3.json > $O/sweep_3.txt 2>&1) &
wait
cp dist_tuto.pth_b200/parallel/allreduce_table.json $O/allreduce_table.json
cat $O/allreduce_table.json
# ---- 4. ResNet-18: per-parameter NCCL | torch DDP | ours | ours bf16 wire | no comm
timeout 900 python bench/resnet_bench.py --gpus 8 --steps 20 --warmup 6 --out $O/resnet_8.json > $O/resnet_8.txt 2>&1; tail -n 1 $O/resnet_8.txt | cut -c1-900
