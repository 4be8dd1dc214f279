#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" > gpurun_out/bn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/bn_tests.log; tail -2 gpurun_out/bn_tests.log
timeout 600 python scripts/gpu_gemm_bn_sweep.py 2>&1 | cut -c1-60
SAM_STAGES=1 timeout 600 python scripts/profile_clip.py 2>&1 | tail -1
