#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or ring" > gpurun_out/r27_tests.log 2>&1; echo "kernel tests rc=$?" >> gpurun_out/r27_tests.log
for m in 2 4 8; do VG_BENCH_M=$m timeout 300 python scripts/gpu_gemv_bench.py; done > gpurun_out/r27_gemv.log 2>&1
timeout 600 python scripts/gpu_batch_decode_time.py > gpurun_out/r27_decode.log 2>&1; echo "rc=$?" >> gpurun_out/r27_decode.log
tail -3 gpurun_out/r27_tests.log; cat gpurun_out/r27_gemv.log; grep "S=\|rc=" gpurun_out/r27_decode.log
