#!/bin/bash
# Cheapest useful 8-GPU call (charged 8x, keep it under ~60 s of box time): the push-exchange test at world 8, then
# bench.py ours at N=8 with the push and the barrier exchange.  Everything else about 8 GPUs is in gpu_8.sh.
set -u
OUT=gpurun_out/multi8q
mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -x -k "push" 2>&1 | tail -12 > $OUT/pytest_push.txt; tail -4 $OUT/pytest_push.txt | cut -c1-200
for push in 1 0; do
  B200DIST_SGD_PUSH=$push timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2962$push \
      bench.py --gpus 8 --steps 2000 --warmup 50 > $OUT/bench_n8_push$push.json 2> $OUT/bench_n8_push$push.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_n8_push$push.json").read().strip().splitlines()[-1])
    print("N=8 push=$push", round(d["value"]), "samples/s", round(d["ms_per_step"] * 1e3, 2), "us/step  e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("N=8 push=$push bench failed", e); print(open("$OUT/bench_n8_push$push.err").read()[-1500:])
PY
done
