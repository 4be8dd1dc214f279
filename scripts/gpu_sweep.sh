#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_gemm_bn_sweep.py > gpurun_out/bn_sweep.log 2>&1; cat gpurun_out/bn_sweep.log | tail -16
