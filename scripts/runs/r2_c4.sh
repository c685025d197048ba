#!/bin/bash
mkdir -p gpurun_out/r2_c4
python scripts/tc_probe_matrix.py --out gpurun_out/r2_c4/matrix.json --only tma4d_inner16_aligned_x0 tma4d_inner16_aligned_x8 tma4d_nhwc_inner32_noswz tma4d_nhwc_inner32_sw32 tma4d_nhwc_inner32_sw32_b1 tma4d_nhwc_oob_batch_zero_fill umma_kmajor_sw32_a umma_mn_major_sw32_a 2>&1 | tee gpurun_out/r2_c4/matrix.txt
