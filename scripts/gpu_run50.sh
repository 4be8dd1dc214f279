#!/bin/bash
mkdir -p gpurun_out
SAM_STAGES=1 timeout 600 python scripts/profile_clip.py 2>&1 | tail -2 > gpurun_out/r50_stages.log
E2E_STAGES=1 timeout 600 python scripts/profile_clip.py 2>&1 | tail -8 >> gpurun_out/r50_stages.log
cat gpurun_out/r50_stages.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file /tmp/r50_launches.csv python scripts/profile_clip.py > gpurun_out/r50_ncu.log 2>&1
python scripts/launches_summarize.py /tmp/r50_launches.csv gpurun_out/r50_launches.md | head -45
