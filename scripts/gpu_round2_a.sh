#!/bin/bash
# GPU call: full test suite with the PARITY report + bench lines for configs 2..5 (each bounded by its own timeout)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rP 2>&1 | grep -E "PARITY|passed|failed|FAILED|Error" > gpurun_out/r02_pytest3.log
tail -3 gpurun_out/r02_pytest3.log
for c in 2 4 3 5; do
  extra="--no-cpu-baseline"; [ "$c" = "2" ] && extra=""
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 $extra > gpurun_out/r02_bench_c$c.json 2> gpurun_out/r02_bench_c$c.err
  echo "config $c rc=$?"; head -c 400 gpurun_out/r02_bench_c$c.json; echo; tail -3 gpurun_out/r02_bench_c$c.err
done
