#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_attn_tc_check.py 2 > gpurun_out/r39_check.log 2>&1; echo "rc=$?" >> gpurun_out/r39_check.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" >> gpurun_out/r39_check.log 2>&1; echo "rc=$?" >> gpurun_out/r39_check.log
tail -15 gpurun_out/r39_check.log
for t in 0 1; do VG_ATTN_TC2=$t timeout 300 python scripts/gpu_attn_tc2_bench.py; done > gpurun_out/r39_bench.log 2>&1
cat gpurun_out/r39_bench.log
