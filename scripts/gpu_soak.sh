#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gpu_soak.py > gpurun_out/soak.log 2>&1; echo "rc=$?" >> gpurun_out/soak.log; tail -6 gpurun_out/soak.log
