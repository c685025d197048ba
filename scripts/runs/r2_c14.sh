#!/bin/bash
mkdir -p gpurun_out/r2_c14
timeout 900 python -m pytest tests -q -m "gpu and not multigpu" --timeout 600 > gpurun_out/r2_c14/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c14/pytest_gpu.txt; tail -n 4 gpurun_out/r2_c14/pytest_gpu.txt
for K in 20 400; do timeout 300 python bench.py --gpus 1 --steps $K --warmup 5 > gpurun_out/r2_c14/n1_k$K.json 2> gpurun_out/r2_c14/n1_k$K.err; python - gpurun_out/r2_c14/n1_k$K.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.3fM us/step %.1f e2e %.3fM"%(d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e6), d['config'].get('e2e_host_us'), (d['config'].get('large_batch') or {}).get('samples_per_s'))
PY
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c14/smoke.txt 2>&1; tail -n 2 gpurun_out/r2_c14/smoke.txt
bash scripts/prof_batched.sh 4096 > gpurun_out/r2_c14/prof.txt 2>&1; tail -n 6 gpurun_out/r2_c14/prof.txt
O=gpurun_out/sanitize; mkdir -p $O
for tool in memcheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_batched.py -q -m gpu -p no:cacheprovider -k "test_every_stage_matches_the_rounding_exact_model and 2-False" > $O/${tool}_batched.txt 2>&1
  echo "$tool batched rc=$?" | tee -a $O/summary.txt; grep -E "ERROR SUMMARY|passed|failed" $O/${tool}_batched.txt | tail -3 | tee -a $O/summary.txt
done
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "test_deterministic_mode_is_bit_reproducible and 32" > $O/memcheck_det.txt 2>&1
echo "memcheck det rc=$?" | tee -a $O/summary.txt; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_det.txt | tail -3 | tee -a $O/summary.txt
