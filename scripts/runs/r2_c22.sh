#!/bin/bash
# call 22 (1 GPU): ncu --set full of the per-sample step kernel (post padded layouts) and the fused all-reduce+SGD kernel, batch 128
O=gpurun_out/ncu_step; mkdir -p $O
timeout 60 ncu --set full --clock-control none --import-source on -k regex:'convnet_step_kernel|allreduce_sgd_kernel' -s 6 -c 2 -f -o $O/step_b128 python scripts/prof_step_once.py fused 128 > $O/step_b128.log 2>&1
echo "ncu rc=$?"
ncu -i $O/step_b128.ncu-rep --page raw --csv > $O/step_b128.raw.csv 2>/dev/null
ls -la $O; tail -3 $O/step_b128.log
