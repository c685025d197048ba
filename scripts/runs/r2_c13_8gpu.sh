#!/bin/bash
O=gpurun_out/r2_c13; mkdir -p $O
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    lb=d['config'].get('large_batch') or {}
    print(sys.argv[1], "N=%d value %.3fM us/step %.1f e2e %s large_batch %s host %s"%(d['n_gpus'], d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value'] and "%.3fM"%(d['e2e']['value']/1e6), lb.get('samples_per_s') and "%.1fM"%(lb['samples_per_s']/1e6), d['config'].get('e2e_host_us')))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8_k20.json 2> $O/n8_k20.err; summ $O/n8_k20.json
(CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --large-batch 0 > $O/n1_k20.json 2> $O/n1_k20.err; summ $O/n1_k20.json) &
(CUDA_VISIBLE_DEVICES=1,2,3,4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus 4 --steps 20 --warmup 5 --large-batch 0 > $O/n4_k20.json 2> $O/n4_k20.err; summ $O/n4_k20.json) &
(CUDA_VISIBLE_DEVICES=5,6 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29803 bench.py --gpus 2 --steps 20 --warmup 5 --large-batch 0 > $O/n2_k20.json 2> $O/n2_k20.err; summ $O/n2_k20.json) &
wait
B200DIST_STRESS_ITERS=100000 timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -s -k "world8_suite_rest" > $O/suite8_rest.txt 2>&1; echo "suite8_rest rc=$?" | tee -a $O/suite8_rest.txt; grep -E "SUITE|passed|failed|Error" $O/suite8_rest.txt | tail -12
