#!/bin/bash
set -u
OUT=gpurun_out/exp2
mkdir -p $OUT
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 "$@" 2>/dev/null | grep "^{" ; }
for pdl in 1 0; do echo "== sgd comm PDL=$pdl"; B200DIST_PDL=$pdl run bench/sgd_comm_bench.py | tee -a $OUT/sgd_comm.jsonl; done
for pdl in 1 0; do for cl in 2 1; do
echo "== bench N=2 PDL=$pdl cluster=$cl"; B200DIST_PDL=$pdl B200DIST_CONVNET_CLUSTER=$cl run bench.py --gpus 2 --steps 400 --warmup 20 --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1e3,2),'us')" | tee -a $OUT/matrix.txt
done; done
