#!/bin/bash
set -u
OUT=gpurun_out/exp3
mkdir -p $OUT
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@" 2>/dev/null | grep "^{" ; }
echo "== sgd comm"; run bench/sgd_comm_bench.py | tee -a $OUT/sgd_comm.jsonl
echo "== bench N=2"; run bench.py --gpus 2 --steps 400 --warmup 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1e3,2),'us e2e', round(d['e2e']['value']))" | tee -a $OUT/matrix.txt
echo "== multi tests"; timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
