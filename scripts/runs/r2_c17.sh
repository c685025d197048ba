#!/bin/bash
# call 17 (1 GPU): verify the cluster-kernel barrier fix (racecheck + memcheck), racecheck the 1-CTA kernel and the batched engine,
# gradient accuracy of every engine vs float64 autograd
set -u
O=gpurun_out/r2_c17; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.txt 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; grep -E '^\{' $O/$name.txt | tail -1 | cut -c1-1200 | tee -a $O/summary.txt; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$name.txt | tail -2 | tee -a $O/summary.txt; }
run racecheck_c4    timeout 300 compute-sanitizer --tool racecheck python scripts/det_diag.py --bsz 32 --steps 3
run memcheck_c4     timeout 300 compute-sanitizer --tool memcheck python scripts/det_diag.py --bsz 32
run racecheck_c2    timeout 300 compute-sanitizer --tool racecheck python scripts/det_diag.py --bsz 64 --steps 2
run racecheck_c1    timeout 300 compute-sanitizer --tool racecheck python scripts/det_diag.py --bsz 128 --steps 2
run memcheck_dettest timeout 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "test_deterministic_mode_is_bit_reproducible"
run racecheck_batched timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_batched.py -q -m gpu -p no:cacheprovider -k "test_every_stage_matches_the_rounding_exact_model[2-False]"
run grads           timeout 200 python scripts/det_diag.py --grads
