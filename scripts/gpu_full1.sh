#!/bin/bash
# Full one-GPU validation: every GPU test that fits one device, smoke, bench (ours + reference).
set -u
OUT=gpurun_out/full1
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > $OUT/pytest.txt; tail -6 $OUT/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -2 $OUT/smoke.txt
timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 50 > $OUT/bench_ours_n1.json 2> $OUT/bench_ours.err; tail -c 600 $OUT/bench_ours_n1.json; tail -3 $OUT/bench_ours.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 300 --warmup 20 > $OUT/bench_ref_n1.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref_n1.json; tail -3 $OUT/bench_ref.err
