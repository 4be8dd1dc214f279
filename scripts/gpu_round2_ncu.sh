#!/bin/bash
# ncu --set full of each hot kernel at its config-2 shape incl. the fused mask decoder (one GPU; no bench number is taken from this run).
# Only the summaries travel back (the .ncu-rep is dropped when it would blow the 64 MiB return limit).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -o gpurun_out/r02_targets -f python scripts/gpu_ncu_targets.py > gpurun_out/r02_ncu_targets.log 2>&1
tail -2 gpurun_out/r02_ncu_targets.log
ncu -i gpurun_out/r02_targets.ncu-rep --page raw --csv > gpurun_out/r02_targets_raw.csv 2>/dev/null
python scripts/ncu_summarize.py gpurun_out/r02_targets_raw.csv gpurun_out/r02_ncu_targets | cut -c1-230
ls -la gpurun_out/
[ $(stat -c %s gpurun_out/r02_targets.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/r02_targets.ncu-rep
gzip -f gpurun_out/r02_targets_raw.csv
