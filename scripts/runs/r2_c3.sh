#!/bin/bash
set -x
mkdir -p gpurun_out/r2_c3
python scripts/tc_probe_matrix.py --out gpurun_out/r2_c3/matrix.json > gpurun_out/r2_c3/matrix.txt 2>&1
python scripts/bt_stage_probe.py 64 > gpurun_out/r2_c3/stages.txt 2>&1
cat gpurun_out/r2_c3/matrix.txt gpurun_out/r2_c3/stages.txt
