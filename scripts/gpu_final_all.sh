#!/bin/bash
# End of round: full GPU suite, smoke(), bench lines of every config, reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rP > gpurun_out/r02_tests_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_final.log; tail -3 gpurun_out/r02_tests_final.log
grep "PARITY" gpurun_out/r02_tests_final.log > gpurun_out/r02_parity_report.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/gpu_final_bench.sh
