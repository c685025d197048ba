#!/bin/bash
# 2 GPUs: every multi-GPU test at world 2, bench N=2 (fused tail on/off), all-reduce sweep (LL variant), ncu attempt on comm kernels
mkdir -p gpurun_out/r2_c10
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc_probe.py tests/test_optim.py -q -m gpu --timeout 600 > gpurun_out/r2_c10/pytest_1gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_c10/pytest_1gpu.txt; tail -n 4 gpurun_out/r2_c10/pytest_1gpu.txt
B200DIST_STRESS_ITERS=20000 timeout 1500 python -m pytest tests/test_gpu_multi.py -q --timeout 900 > gpurun_out/r2_c10/pytest_multi2.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_c10/pytest_multi2.txt; tail -n 12 gpurun_out/r2_c10/pytest_multi2.txt
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps ${STEPS:-400} --warmup 5 --large-batch ${LB:-0} > gpurun_out/r2_c10/$name.json 2> gpurun_out/r2_c10/$name.err; python - gpurun_out/r2_c10/$name.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('large_batch'))
PY
}
STEPS=20 LB=4096 run n2_k20
run n2_k400
run n2_k400_fusedtail B200DIST_FUSED_TAIL=1
timeout 600 python bench/allreduce_sweep.py --gpus 2 --max-mb 64 --out gpurun_out/r2_c10/sweep_2.json > gpurun_out/r2_c10/sweep.txt 2>&1; tail -n 3 gpurun_out/r2_c10/sweep.txt | cut -c1-600
bash scripts/prof_comm.sh 2 > gpurun_out/r2_c10/prof_comm.txt 2>&1; tail -n 8 gpurun_out/r2_c10/prof_comm.txt
