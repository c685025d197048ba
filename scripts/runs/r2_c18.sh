#!/bin/bash
# call 18 (2 GPUs): trainer-vs-torch tests with fp32 references, NVLink byte counters per all-reduce variant, N=2 bench
set -u
O=gpurun_out/r2_c18; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider -k "test_fused_trainer_equals_global_batch_sgd or test_batched_tensor_core_trainer or test_bf16_wire" > $O/pytest_multi2.txt 2>&1
echo "multi2 rc=$?" | tee -a $O/summary.txt; tail -3 $O/pytest_multi2.txt | tee -a $O/summary.txt
timeout 240 python bench/nvlink_bytes.py --gpus 2 --max-mb 64 --out $O/nvlink_bytes_2.json > $O/nvlink.txt 2>&1
echo "nvlink rc=$?" | tee -a $O/summary.txt; grep -E '^\{' $O/nvlink.txt | cut -c1-400 | tee -a $O/summary.txt; tail -5 $O/nvlink.txt | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --large-batch 0 > $O/n2_k20.json 2> $O/n2_k20.err
echo "bench rc=$?" | tee -a $O/summary.txt; grep -E '^\{' $O/n2_k20.json | cut -c1-600 | tee -a $O/summary.txt
