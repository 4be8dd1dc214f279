#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/pre_tests.log 2>&1; echo "rc=$?" >> gpurun_out/pre_tests.log; tail -12 gpurun_out/pre_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/pre_bench_c2.json 2> gpurun_out/pre_bench_c2.err; tail -c 300 gpurun_out/pre_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/pre_bench_c2.json').read().strip().splitlines()[-1])
print(round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('passes_ms'), 'h2d', d['e2e']['h2d_bytes_per_step'], 'fp32', d['e2e']['with_fp32_sam_inputs']['value'])
PY
