#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sam2_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/elem_tests.log 2>&1; echo "rc=$?" >> gpurun_out/elem_tests.log; tail -3 gpurun_out/elem_tests.log
bash scripts/gpu_ncu_elementwise.sh | tail -11
