#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r42_check.log 2>&1; echo "rc=$?" >> gpurun_out/r42_check.log
tail -3 gpurun_out/r42_check.log
VG_ATTN_TC2=1 timeout 300 python scripts/gpu_attn_tc2_bench.py > gpurun_out/r42_bench.log 2>&1
cat gpurun_out/r42_bench.log
