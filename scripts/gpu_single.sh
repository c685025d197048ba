#!/bin/bash
# One-GPU measurement pass: kernel numerics, smoke, bench (both arms), launch list + ncu capture.
# Everything is bounded by `timeout`; outputs land in gpurun_out/.
set -u
OUT=gpurun_out/single
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvsmi.txt 2>&1
echo "== tests" ; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > $OUT/pytest.txt; tail -15 $OUT/pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -3 $OUT/smoke.txt
echo "== bench ours"; timeout 600 python bench.py --gpus 1 --steps 400 --warmup 20 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; tail -c 1500 $OUT/bench_ours.json; tail -5 $OUT/bench_ours.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --gpus 1 --steps 400 --warmup 20 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 1500 $OUT/bench_ref.json; tail -5 $OUT/bench_ref.err
echo "== kernel microbench"; timeout 600 python bench/kernel_bench.py > $OUT/kernel_bench.json 2> $OUT/kernel_bench.err; tail -c 3000 $OUT/kernel_bench.json; tail -5 $OUT/kernel_bench.err
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"convnet_step|allreduce_sgd" -s 10 -c 40 --csv --log-file $OUT/launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --graph-chunk 1 --no-e2e > $OUT/ncu_launch.log 2>&1
tail -3 $OUT/launches.csv
echo "== ncu full (convnet_step)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:convnet_step -s 6 -c 2 -o $OUT/prof_convnet -f python bench.py --gpus 1 --steps 8 --warmup 3 --graph-chunk 1 --no-e2e > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 1 -o $OUT/prof_gemm -f python bench/kernel_bench.py --gemm-only --big-only > $OUT/ncu_gemm.log 2>&1
ls -la $OUT
