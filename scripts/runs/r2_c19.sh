#!/bin/bash
# call 19 (1 GPU): flag-mode executor (no cross-stream events): tests + e2e with flags on / off
set -u
O=gpurun_out/r2_c19; mkdir -p $O
timeout 600 python -m pytest tests -q -m "gpu and not multigpu" -p no:cacheprovider -x > $O/pytest_gpu.txt 2>&1
echo "pytest rc=$?" | tee -a $O/summary.txt; tail -4 $O/pytest_gpu.txt | tee -a $O/summary.txt
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'value %.3fM us/step %.1f e2e %.3fM' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('e2e_host_us'), d['clocks']['sm_mhz'], (d['config'].get('large_batch') or {}).get('value'))
PY
}
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/n1_k20_flags.json 2> $O/n1_k20_flags.err; show $O/n1_k20_flags.json | tee -a $O/summary.txt
B200DIST_EXEC_FLAGS=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --large-batch 0 > $O/n1_k20_events.json 2> $O/n1_k20_events.err; show $O/n1_k20_events.json | tee -a $O/summary.txt
timeout 300 python bench.py --gpus 1 --steps 400 --warmup 20 --large-batch 0 > $O/n1_k400_flags.json 2> $O/n1_k400_flags.err; show $O/n1_k400_flags.json | tee -a $O/summary.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --large-batch 0 > $O/n1_k20_flags_b.json 2> $O/n1_k20_flags_b.err; show $O/n1_k20_flags_b.json | tee -a $O/summary.txt
tail -3 $O/*.err | cut -c1-300
