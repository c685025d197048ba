#!/bin/bash
# NVLink-level ncu capture of the peer-memory comm kernels at N GPUs (application replay: every pass re-runs all ranks).
set -x
N=${1:-2}
OUT=gpurun_out/ncu_comm_n$N; mkdir -p $OUT
timeout 900 ncu --replay-mode application --target-processes all --clock-control none \
  --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
  -k regex:'allreduce_.*kernel' -f -o $OUT/comm python scripts/prof_comm_worker.py $N > $OUT/ncu.log 2>&1
echo "ncu rc=$?" >> $OUT/ncu.log
tail -n 20 $OUT/ncu.log
ncu -i $OUT/comm.ncu-rep --page raw --csv > $OUT/comm.raw.csv 2>/dev/null
ncu -i $OUT/comm.ncu-rep --page details --csv > $OUT/comm.details.csv 2>/dev/null
ls -la $OUT
