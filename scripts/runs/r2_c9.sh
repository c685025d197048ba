#!/bin/bash
mkdir -p gpurun_out/r2_c9
timeout 1500 python -m pytest tests -q -m "gpu and not multigpu" --timeout 900 > gpurun_out/r2_c9/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c9/pytest_gpu.txt
tail -n 6 gpurun_out/r2_c9/pytest_gpu.txt
run() { name=$1; shift; env "$@" timeout 300 python bench.py --gpus 1 --steps ${STEPS:-400} --warmup 5 --large-batch 0 > gpurun_out/r2_c9/$name.json 2> gpurun_out/r2_c9/$name.err; python - gpurun_out/r2_c9/$name.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('e2e_host_us'))
PY
}
STEPS=20 run k20_default
STEPS=20 run k20_default_b
STEPS=400
run k400_default
run k400_graphmode B200DIST_EXEC_DIRECT=0
run k400_fusedtail B200DIST_FUSED_TAIL=1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_c9/full_k20.json 2> gpurun_out/r2_c9/full_k20.err; tail -c 1500 gpurun_out/r2_c9/full_k20.json
