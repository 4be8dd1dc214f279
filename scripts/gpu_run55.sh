#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_towers_batch.py > gpurun_out/r55.log 2>&1; tail -5 gpurun_out/r55.log
