#!/bin/bash
mkdir -p gpurun_out/r2_c8
timeout 1500 python -m pytest tests -q -m "gpu and not multigpu" --timeout 900 > gpurun_out/r2_c8/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c8/pytest_gpu.txt
tail -n 8 gpurun_out/r2_c8/pytest_gpu.txt
timeout 600 python bench/batched_bench.py --batch 4096 --steps 20 --no-simt --out gpurun_out/r2_c8/batched_bench.json > gpurun_out/r2_c8/bench.txt 2>&1; tail -n 2 gpurun_out/r2_c8/bench.txt
run() { name=$1; shift; env "$@" timeout 300 python bench.py --gpus 1 --steps ${STEPS:-400} --warmup 5 --large-batch 0 > gpurun_out/r2_c8/$name.json 2> gpurun_out/r2_c8/$name.err; python - gpurun_out/r2_c8/$name.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('e2e_host_us'))
PY
}
STEPS=20 run k20 B200DIST_FUSED_TAIL=0
STEPS=400
run base_fused0 B200DIST_FUSED_TAIL=0
run thr1 B200DIST_FUSED_TAIL=0 B200DIST_LOADER_THREADS=1
run chunk1 B200DIST_FUSED_TAIL=0 B200DIST_EXEC_CHUNK=1 B200DIST_LOADER_THREADS=1
run chunk4 B200DIST_FUSED_TAIL=0 B200DIST_EXEC_CHUNK=4 B200DIST_LOADER_THREADS=1
run thr4 B200DIST_FUSED_TAIL=0 B200DIST_LOADER_THREADS=4
nproc; python -c "import os; print(os.cpu_count(), len(os.sched_getaffinity(0)))"; cat /sys/fs/cgroup/cpu.max 2>/dev/null
