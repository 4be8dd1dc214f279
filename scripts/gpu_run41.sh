#!/bin/bash
mkdir -p gpurun_out
for k in 2 3; do VG_ATTN_KST=$k VG_ATTN_TC2=1 timeout 300 python scripts/gpu_attn_tc2_bench.py 2>&1 | sed "s/^/KST=$k /"; done > gpurun_out/r41_bench.log 2>&1
cat gpurun_out/r41_bench.log
