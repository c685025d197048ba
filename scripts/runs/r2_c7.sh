#!/bin/bash
mkdir -p gpurun_out/r2_c7
timeout 1500 python -m pytest tests -q -m "gpu and not multigpu" --timeout 900 -x > gpurun_out/r2_c7/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c7/pytest_gpu.txt
tail -n 15 gpurun_out/r2_c7/pytest_gpu.txt
timeout 600 python bench/batched_bench.py --batch 1024 4096 16384 --steps 20 --out gpurun_out/r2_c7/batched_bench.json > gpurun_out/r2_c7/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/r2_c7/bench.txt
tail -n 6 gpurun_out/r2_c7/bench.txt
for K in 20 400; do timeout 300 python bench.py --gpus 1 --steps $K --warmup 5 --large-batch 0 > gpurun_out/r2_c7/n1_k$K.json 2> gpurun_out/r2_c7/n1_k$K.err; python - gpurun_out/r2_c7/n1_k$K.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM launches %d"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6, d['gpu_launches']))
PY
tail -n 3 gpurun_out/r2_c7/n1_k$K.err; done
B200DIST_FUSED_TAIL=0 timeout 300 python bench.py --gpus 1 --steps 400 --warmup 5 --large-batch 0 > gpurun_out/r2_c7/n1_k400_2kernel.json 2>/dev/null; python - gpurun_out/r2_c7/n1_k400_2kernel.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6))
PY
