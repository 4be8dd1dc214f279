#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention or pool" > gpurun_out/r46_check.log 2>&1; echo "rc=$?" >> gpurun_out/r46_check.log
tail -3 gpurun_out/r46_check.log
for pz in 0 1; do VG_ATTN_PERSIST=$pz VG_ATTN_TC2=0 timeout 300 python scripts/gpu_attn_tc2_bench.py 2>&1 | sed "s/^/PERSIST=$pz /"; done > gpurun_out/r46_bench.log 2>&1
cat gpurun_out/r46_bench.log
VG_ATTN_PERSIST=1 timeout 300 python scripts/gpu_attn_tc_check.py 2>&1 | grep "bench" | grep "legacy=False" | head -1
