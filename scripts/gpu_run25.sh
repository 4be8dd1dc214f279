#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or ring" > gpurun_out/r25_tests.log 2>&1; echo "kernel tests rc=$?" >> gpurun_out/r25_tests.log
timeout 600 python scripts/gpu_batch_decode_time.py > gpurun_out/r25_decode.log 2>&1; echo "rc=$?" >> gpurun_out/r25_decode.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py -x -q > gpurun_out/r25_pipe.log 2>&1; echo "rc=$?" >> gpurun_out/r25_pipe.log
tail -5 gpurun_out/r25_tests.log; grep "S=\|rc=" gpurun_out/r25_decode.log; tail -3 gpurun_out/r25_pipe.log
