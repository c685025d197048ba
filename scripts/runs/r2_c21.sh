#!/bin/bash
# call 21 (1 GPU): final single-GPU tier on the final tree + smoke + driver-contract bench
set -u
O=gpurun_out/r2_c21; mkdir -p $O
timeout 400 python -m pytest tests -q -m "gpu and not multigpu" -p no:cacheprovider > $O/pytest_gpu.txt 2>&1
echo "pytest rc=$?" | tee -a $O/summary.txt; grep -E "passed|failed|^FAILED|^ERROR|Error" $O/pytest_gpu.txt | tail -12 | cut -c1-400 | tee -a $O/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -2 $O/smoke.txt | cut -c1-300 | tee -a $O/summary.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/n1_k20.json 2> $O/n1_k20.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'PY' | tee -a gpurun_out/r2_c21/summary.txt
import json
for l in open('gpurun_out/r2_c21/n1_k20.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.3fM us/step %.1f e2e %.3fM' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('e2e_host_us'), d['clocks']['sm_mhz'], (d['config'].get('large_batch') or {}))
PY
