#!/bin/bash
mkdir -p gpurun_out
for m in 1 2 4; do VG_BENCH_M=$m timeout 300 python scripts/gpu_gemv_bench.py; done > gpurun_out/r26_gemv.log 2>&1
VG_BENCH_M=4 VG_BENCH_REPS=1 timeout 600 ncu -k regex:gemv_ring_mr -c 6 --set full --import-source on --clock-control none -o /tmp/r26 python scripts/gpu_gemv_bench.py > gpurun_out/r26_ncu.log 2>&1
ncu -i /tmp/r26.ncu-rep --page source --csv > gpurun_out/r26_source.csv 2>/dev/null
ncu -i /tmp/r26.ncu-rep --page raw --csv > gpurun_out/r26_raw.csv 2>/dev/null
cat gpurun_out/r26_gemv.log; ls -la gpurun_out/r26*
