#!/bin/bash
# final 1-GPU call: ncu evidence for the tensor-core training kernels + sanitizer pass on the new kernels
mkdir -p gpurun_out/r2_c12
bash scripts/prof_batched.sh 4096 > gpurun_out/r2_c12/prof.txt 2>&1; tail -n 8 gpurun_out/r2_c12/prof.txt
O=gpurun_out/sanitize; mkdir -p $O
for tool in memcheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_batched.py -q -m gpu -p no:cacheprovider -k "test_every_stage_matches_the_rounding_exact_model and 2-False" > $O/${tool}_batched.txt 2>&1
  echo "$tool batched rc=$?" | tee -a $O/summary.txt; grep -E "ERROR SUMMARY|passed|failed" $O/${tool}_batched.txt | tail -3 | tee -a $O/summary.txt
done
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "test_deterministic_mode_is_bit_reproducible and 32" > $O/memcheck_det.txt 2>&1
echo "memcheck det rc=$?" | tee -a $O/summary.txt; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_det.txt | tail -3 | tee -a $O/summary.txt
