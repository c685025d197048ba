#!/bin/bash
# Two-GPU validation of the push exchange: multi-GPU tests, then bench.py at N=2 with the barrier and the push exchange.
set -u
OUT=gpurun_out/push2
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "rewinds" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -30 > $OUT/pytest_multi.txt; tail -12 $OUT/pytest_multi.txt | cut -c1-200
for push in 0 1; do
  B200DIST_SGD_PUSH=$push timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2000 --warmup 50 > $OUT/bench_n2_push$push.json 2> $OUT/bench_n2_push$push.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_n2_push$push.json").read().strip().splitlines()[-1])
    print("push=$push", round(d["value"]), "samples/s", round(d["ms_per_step"] * 1e3, 2), "us/step  e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("push=$push bench failed", e); print(open("$OUT/bench_n2_push$push.err").read()[-1500:])
PY
done
