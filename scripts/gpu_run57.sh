#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sam2_gpu.py -x -q > gpurun_out/r57_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r57_tests.log; tail -4 gpurun_out/r57_tests.log
SAM_STAGES=1 timeout 600 python scripts/profile_clip.py 2>&1 | tail -1
