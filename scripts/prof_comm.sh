#!/bin/bash
# NVLink byte counters of the peer-memory comm kernels at N GPUs.  Only rank 0 runs under ncu, with metrics that need ONE pass
# (no kernel replay: the kernels spin on their peers, which run unprofiled and would not replay with them).
set -x
N=${1:-2}
OUT=gpurun_out/ncu_comm_n$N; mkdir -p $OUT
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlink" | head -60 > $OUT/nvlink_metrics_available.txt
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29791 WORLD_SIZE=$N
for r in $(seq 1 $((N-1))); do RANK=$r LOCAL_RANK=$r timeout 600 python scripts/prof_comm_rank.py > $OUT/rank$r.log 2>&1 & done
M="nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,sm__cycles_active.avg,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum"
RANK=0 LOCAL_RANK=0 timeout 600 ncu --metrics $M --clock-control none --cache-control none --replay-mode kernel \
  -k regex:'allreduce_.*kernel|convnet_' -f -o $OUT/comm_rank0 --csv --log-file $OUT/ncu_rank0.log python scripts/prof_comm_rank.py > $OUT/rank0.log 2>&1
echo "ncu rc=$?" >> $OUT/rank0.log
wait
tail -n 5 $OUT/rank0.log $OUT/rank1.log
ncu -i $OUT/comm_rank0.ncu-rep --page raw --csv > $OUT/comm_rank0.raw.csv 2>/dev/null
head -c 3000 $OUT/comm_rank0.raw.csv
ls -la $OUT
