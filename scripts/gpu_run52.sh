#!/bin/bash
mkdir -p gpurun_out
VG_GEMM_SHAPES=1 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r52.json 2> gpurun_out/r52.err
grep "gemm shape" gpurun_out/r52.err | head -40
