#!/bin/bash
# usage (under gpurun): bash scripts/gpu_run_tests.sh [pytest args]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
