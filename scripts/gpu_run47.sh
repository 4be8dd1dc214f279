#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_attn_tc2_bench.py > gpurun_out/r47_bench.log 2>&1; cat gpurun_out/r47_bench.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r47_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r47_tests.log; tail -3 gpurun_out/r47_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r47_bench_c2.json 2> gpurun_out/r47_bench_c2.err; tail -c 300 gpurun_out/r47_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r47_bench_c2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('value_decode_batch_1'), d['config']['latency_ms_single_clip'], d['roofline']['frac'], d['clocks'])
PY
