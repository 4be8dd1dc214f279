#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu -k regex:gemm_tcgen05 -c 2 --set full --import-source on --clock-control none -o /tmp/r58 python scripts/gpu_gemm_one.py > gpurun_out/r58_ncu.log 2>&1
ncu -i /tmp/r58.ncu-rep --page source --csv > gpurun_out/r58_source.csv 2>/dev/null
ncu -i /tmp/r58.ncu-rep --page raw --csv > gpurun_out/r58_raw.csv 2>/dev/null
ls -la gpurun_out/r58*
