#!/bin/bash
mkdir -p gpurun_out
VG_BENCH_SHAPE=4 VG_BENCH_M=4 VG_BENCH_REPS=1 timeout 600 ncu -k regex:gemv_ring_mma -c 2 --set full --import-source on --clock-control none -o /tmp/r31 python scripts/gpu_gemv_bench.py > gpurun_out/r31_ncu.log 2>&1
ncu -i /tmp/r31.ncu-rep --page source --csv > gpurun_out/r31_source.csv 2>/dev/null
ncu -i /tmp/r31.ncu-rep --page raw --csv > gpurun_out/r31_raw.csv 2>/dev/null
ls -la gpurun_out/r31*
