#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_lmm_stage_time.py 12 > gpurun_out/r35.log 2>&1; tail -3 gpurun_out/r35.log
