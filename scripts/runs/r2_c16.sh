#!/bin/bash
# call 16 (1 GPU): where does deterministic mode diverge under compute-sanitizer (4-CTA clusters)?  + TF32 noise of the torch reference
set -u
O=gpurun_out/r2_c16; mkdir -p $O
run() { name=$1; shift; ( "$@" ) > $O/$name.txt 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; grep -E '^\{' $O/$name.txt | tail -1 | cut -c1-1500 | tee -a $O/summary.txt; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/$name.txt | tail -1 | tee -a $O/summary.txt; }
run plain_c4        timeout 200 python scripts/det_diag.py --bsz 32
run memcheck_c4     timeout 300 compute-sanitizer --tool memcheck python scripts/det_diag.py --bsz 32
run memcheck_c4_eager timeout 300 compute-sanitizer --tool memcheck python scripts/det_diag.py --bsz 32 --eager
run memcheck_c4_nopdl env B200DIST_PDL=0 timeout 300 compute-sanitizer --tool memcheck python scripts/det_diag.py --bsz 32
run memcheck_c2     timeout 300 compute-sanitizer --tool memcheck python scripts/det_diag.py --bsz 32 --cluster 2
run racecheck_c4    timeout 300 compute-sanitizer --tool racecheck python scripts/det_diag.py --bsz 32 --steps 3
run synccheck_c4    timeout 300 compute-sanitizer --tool synccheck python scripts/det_diag.py --bsz 32 --steps 3
run tf32            timeout 200 python scripts/det_diag.py --tf32
