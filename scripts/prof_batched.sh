#!/bin/bash
# ncu captures of the batched tensor-core engine's kernels (1 GPU).  Reports + raw-page CSVs land in gpurun_out/ncu_batched/.
set -x
OUT=gpurun_out/ncu_batched; mkdir -p $OUT
B=${1:-4096}
KERNELS=${2:-"bt_conv2_fwd bt_conv2_dgrad bt_conv2_wgrad"}
# every launch of 2 steps with its device time (cold-cache, serialised: compare SHARES)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file $OUT/launches_B$B.csv python scripts/prof_step_once.py batched $B > $OUT/launches.log 2>&1
for K in $KERNELS; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/$K python scripts/prof_step_once.py batched $B > $OUT/$K.log 2>&1
  ncu -i $OUT/$K.ncu-rep --page raw --csv > $OUT/$K.raw.csv 2>/dev/null
  ncu -i $OUT/$K.ncu-rep --page details --csv > $OUT/$K.details.csv 2>/dev/null
done
ls -la $OUT
