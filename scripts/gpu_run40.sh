#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu -k regex:flash_attn_tc2 -c 1 --set full --import-source on --clock-control none -o /tmp/r40 python scripts/gpu_attn_one.py > gpurun_out/r40_ncu.log 2>&1
ncu -i /tmp/r40.ncu-rep --page source --csv > gpurun_out/r40_source.csv 2>/dev/null
ncu -i /tmp/r40.ncu-rep --page raw --csv > gpurun_out/r40_raw.csv 2>/dev/null
ls -la gpurun_out/r40*
