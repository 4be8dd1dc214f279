#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" > gpurun_out/r44_check.log 2>&1; echo "rc=$?" >> gpurun_out/r44_check.log
tail -3 gpurun_out/r44_check.log
timeout 300 python scripts/gpu_gemm_time.py > gpurun_out/r44_gemm.log 2>&1; cat gpurun_out/r44_gemm.log
