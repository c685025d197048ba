#!/bin/bash
mkdir -p gpurun_out/r2_c6
python scripts/tc_probe_matrix.py --out gpurun_out/r2_c6/matrix.json --only umma_window_fwd_kmajor_sw32 umma_window_wgrad_mnmajor_sw32 tma4d_window_image_y_b_x_c 2>&1 | tee gpurun_out/r2_c6/matrix.txt
timeout 900 python -m pytest tests/test_gpu_batched.py -q --timeout 600 > gpurun_out/r2_c6/batched.txt 2>&1; echo "batched rc=$?" >> gpurun_out/r2_c6/batched.txt
tail -n 12 gpurun_out/r2_c6/batched.txt
cp gpurun_out/batched_diag_*.json gpurun_out/r2_c6/ 2>/dev/null
timeout 600 python bench/batched_bench.py --batch 1024 4096 16384 --steps 20 --out gpurun_out/r2_c6/batched_bench.json > gpurun_out/r2_c6/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/r2_c6/bench.txt
tail -n 12 gpurun_out/r2_c6/bench.txt
for K in 20 400; do timeout 300 python bench.py --gpus 1 --steps $K --warmup 5 --large-batch 0 > gpurun_out/r2_c6/n1_k$K.json 2> gpurun_out/r2_c6/n1_k$K.err; python - gpurun_out/r2_c6/n1_k$K.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6))
PY
tail -n 3 gpurun_out/r2_c6/n1_k$K.err; done
B200DIST_LOADER_THREADS=1 timeout 300 python bench.py --gpus 1 --steps 400 --warmup 5 --large-batch 0 > gpurun_out/r2_c6/n1_k400_1thread.json 2>/dev/null; python - gpurun_out/r2_c6/n1_k400_1thread.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6))
PY
