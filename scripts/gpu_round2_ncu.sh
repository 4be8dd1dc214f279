#!/bin/bash
# ncu captures (one GPU): (1) --set full of each hot kernel at its config-2 shape incl. the fused mask decoder, (2) launch list of one config-2 clip
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -o gpurun_out/r02_targets -f python scripts/gpu_ncu_targets.py > gpurun_out/r02_ncu_targets.log 2>&1
ncu -i gpurun_out/r02_targets.ncu-rep --page raw --csv > gpurun_out/r02_targets_raw.csv 2>/dev/null
python scripts/ncu_summarize.py gpurun_out/r02_targets_raw.csv gpurun_out/r02_ncu_targets | cut -c1-220
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 20000 --csv --log-file gpurun_out/r02_launches.csv python scripts/profile_clip.py > gpurun_out/r02_prof.log 2>&1
tail -3 gpurun_out/r02_prof.log
python scripts/launches_summarize.py gpurun_out/r02_launches.csv gpurun_out/r02_launches.md | head -40
