#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_pipeline_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/r33_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r33_tests.log
tail -4 gpurun_out/r33_tests.log
timeout 900 python bench.py --config 3 > gpurun_out/r33_bench_c3.json 2> gpurun_out/r33_bench_c3.err; tail -1 gpurun_out/r33_bench_c3.json
