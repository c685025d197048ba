#!/bin/bash
# call 20 (2 GPUs): e2e with the common start instant (both arms)
set -u
O=gpurun_out/r2_c20; mkdir -p $O
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d.get('impl'), 'value %.3fM us/step %.1f e2e %.3fM' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('e2e_host_us'), d['clocks']['sm_mhz'])
PY
}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --large-batch 0 > $O/n2_k20.json 2> $O/n2_k20.err; echo "ours rc=$?" | tee -a $O/summary.txt; show $O/n2_k20.json | tee -a $O/summary.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > $O/ref_n2.json 2> $O/ref_n2.err; echo "ref rc=$?" | tee -a $O/summary.txt; show $O/ref_n2.json | tee -a $O/summary.txt
tail -n 3 $O/n2_k20.err $O/ref_n2.err | cut -c1-300
