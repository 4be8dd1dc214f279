#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_pipe_trace.py > gpurun_out/r36.log 2>&1; grep -v "^\[" gpurun_out/r36.log | tail -60
