#!/bin/bash
mkdir -p gpurun_out
VG_ATTN_TC2=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r43_check.log 2>&1; echo "rc=$?" >> gpurun_out/r43_check.log
tail -3 gpurun_out/r43_check.log
VG_ATTN_TC2=0 timeout 300 python scripts/gpu_attn_tc2_bench.py > gpurun_out/r43_bench.log 2>&1
cat gpurun_out/r43_bench.log
VG_ATTN_TC2=0 timeout 300 python scripts/gpu_attn_tc_check.py 2>&1 | grep bench | grep "legacy=False" | head -3
