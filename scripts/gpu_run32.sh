#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or ring" > gpurun_out/r32_tests.log 2>&1; echo "kernel tests rc=$?" >> gpurun_out/r32_tests.log
tail -3 gpurun_out/r32_tests.log
for seg in 2048 1024; do echo "SEG=$seg"; VG_GEMV_RING_MR_SEG=$seg VG_BENCH_M=4 timeout 300 python scripts/gpu_gemv_bench.py; done > gpurun_out/r32_sweep.log 2>&1
VG_BATCH_TC_MIN=9 timeout 600 python scripts/gpu_batch_decode_time.py 2>&1 | grep "S=" >> gpurun_out/r32_sweep.log
cat gpurun_out/r32_sweep.log
