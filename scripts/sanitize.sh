#!/bin/bash
# Race / memory checking of the hand-written kernels (run on a GPU box; slow -- small shapes only).
#   memcheck : out-of-bounds / misaligned accesses in every kernel of the single-GPU test tier
#   racecheck: shared-memory hazards in the fused ConvNet kernel (phase barriers) and the all-reduce/SGD kernel
#   synccheck: divergent barriers
set -u
OUT=gpurun_out/sanitize
mkdir -p $OUT
SEL='test_convnet_loss_and_grads_match_autograd[16] or test_convnet_training_dropout or test_sgd_flat or test_tcgen05_gemm_matches_torch[128-64-64]'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "$SEL" > $OUT/$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/$tool.txt | tail -3
done
