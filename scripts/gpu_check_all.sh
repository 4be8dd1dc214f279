#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/all_tests.log; tail -4 gpurun_out/all_tests.log
timeout 900 python bench.py --config 4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/all_bench_c4.json 2> gpurun_out/all_bench_c4.err; tail -c 200 gpurun_out/all_bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/all_bench_c4.json').read().strip().splitlines()[-1])
print(round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('passes_ms'), 'db1', d['config'].get('value_decode_batch_1'))
PY
