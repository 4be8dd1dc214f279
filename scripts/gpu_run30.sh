#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or ring" > gpurun_out/r30_tests.log 2>&1; echo "kernel tests rc=$?" >> gpurun_out/r30_tests.log
tail -3 gpurun_out/r30_tests.log
for kb in 224 112; do echo "SMEM_KB=$kb"; VG_GEMV_RING_MR_SMEM_KB=$kb VG_BENCH_M=4 timeout 300 python scripts/gpu_gemv_bench.py; VG_GEMV_RING_MR_SMEM_KB=$kb VG_BATCH_TC_MIN=9 timeout 600 python scripts/gpu_batch_decode_time.py 2>&1 | grep "S="; done > gpurun_out/r30_sweep.log 2>&1
cat gpurun_out/r30_sweep.log
