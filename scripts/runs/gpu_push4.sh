#!/bin/bash
# Four-GPU check of the push exchange (world > 2: peer rotation, inbox indexing) + bench at N=4.
set -u
OUT=gpurun_out/push4
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -x -k "push or fused_trainer" 2>&1 | tail -12 > $OUT/pytest_multi.txt; tail -6 $OUT/pytest_multi.txt | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 2000 --warmup 50 > $OUT/bench_n4_push1.json 2> $OUT/bench_n4_push1.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_n4_push1.json").read().strip().splitlines()[-1])
    print("N=4 push", round(d["value"]), "samples/s", round(d["ms_per_step"] * 1e3, 2), "us/step  e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_n4_push1.err").read()[-1500:])
PY
