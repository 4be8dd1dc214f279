#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rP > gpurun_out/r48_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r48_tests.log; tail -3 gpurun_out/r48_tests.log
grep "PARITY" gpurun_out/r48_tests.log > gpurun_out/r48_parity.log; wc -l gpurun_out/r48_parity.log
