#!/bin/bash
mkdir -p gpurun_out
VG_BENCH_M=4 VG_BENCH_REPS=1 timeout 600 ncu -k regex:gemv_ring_mr -c 2 --set full --import-source on --clock-control none -o /tmp/r28 python scripts/gpu_gemv_bench.py > gpurun_out/r28_ncu.log 2>&1
ncu -i /tmp/r28.ncu-rep --page source --csv > gpurun_out/r28_source.csv 2>/dev/null
ncu -i /tmp/r28.ncu-rep --page raw --csv > gpurun_out/r28_raw.csv 2>/dev/null
ls -la gpurun_out/r28*
