#!/bin/bash
O=gpurun_out/r2_c15; mkdir -p $O
# --- determinism under perturbed timing (1 GPU)
T='tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k test_deterministic_mode_is_bit_reproducible'
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest $T > $O/det_launch_blocking.txt 2>&1; echo "det CUDA_LAUNCH_BLOCKING rc=$?" | tee -a $O/det_summary.txt
B200DIST_PDL=0 timeout 300 python -m pytest $T > $O/det_nopdl.txt 2>&1; echo "det PDL=0 rc=$?" | tee -a $O/det_summary.txt
B200DIST_CONVNET_CLUSTER=1 timeout 400 compute-sanitizer --tool memcheck python -m pytest $T > $O/det_memcheck_cluster1.txt 2>&1; echo "det memcheck cluster=1 rc=$?" | tee -a $O/det_summary.txt
timeout 400 compute-sanitizer --tool memcheck python -m pytest $T > $O/det_memcheck_default.txt 2>&1; echo "det memcheck default rc=$?" | tee -a $O/det_summary.txt
grep -h -E "passed|failed" $O/det_*.txt | tee -a $O/det_summary.txt
# --- 2 GPUs: final code through the whole multi-GPU tier
B200DIST_STRESS_ITERS=20000 timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 > $O/pytest_multi2.txt 2>&1; echo "multi2 rc=$?" | tee -a $O/pytest_multi2.txt; tail -n 4 $O/pytest_multi2.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 --large-batch 0 > $O/n2_k20.json 2> $O/n2_k20.err
python - $O/n2_k20.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('e2e_host_us'), d['clocks'])
PY
bash scripts/prof_comm.sh 2 > $O/prof_comm.txt 2>&1; tail -n 25 $O/prof_comm.txt | cut -c1-400
