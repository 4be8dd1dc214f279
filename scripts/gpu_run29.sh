#!/bin/bash
mkdir -p gpurun_out
for pw in 1 2 8; do for seg in 1024 2048 4096; do echo "PW=$pw SEG=$seg"; VG_GEMV_RING_MR_PW=$pw VG_GEMV_RING_MR_SEG=$seg VG_BENCH_M=4 timeout 300 python scripts/gpu_gemv_bench.py; done; done > gpurun_out/r29_gemv.log 2>&1
cat gpurun_out/r29_gemv.log
