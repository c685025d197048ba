#!/bin/bash
# One-GPU part of the push-exchange validation: emulated two-rank protocol test, state-dict rewind, executor tests.
set -u
OUT=gpurun_out/push1
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "push or rewinds or executor or fused_trainer" 2>&1 | tail -30 > $OUT/pytest.txt; tail -30 $OUT/pytest.txt | cut -c1-200
