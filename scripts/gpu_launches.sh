#!/bin/bash
# ncu launch list of one config-2 clip (per-launch gpu__time_duration), summarised on the box; stage timings beside it.
mkdir -p gpurun_out
E2E_STAGES=1 timeout 600 python scripts/profile_clip.py 2>&1 | tail -7 > gpurun_out/r02_stages.log
SAM_STAGES=1 timeout 600 python scripts/profile_clip.py 2>&1 | tail -1 >> gpurun_out/r02_stages.log
cat gpurun_out/r02_stages.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file /tmp/r02_launches.csv python scripts/profile_clip.py > gpurun_out/r02_launches_ncu.log 2>&1
python scripts/launches_summarize.py /tmp/r02_launches.csv gpurun_out/r02_launches.md | head -30
