#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pipeline_gpu.py tests/test_e2e_gpu.py tests/test_fullsize_gpu.py tests/test_preprocess_gpu.py -x -q > gpurun_out/lmm_tests.log 2>&1; echo "rc=$?" >> gpurun_out/lmm_tests.log; tail -4 gpurun_out/lmm_tests.log
timeout 900 python scripts/gpu_soak.py 2>&1 | grep "^pass"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/lmm_bench_c2.json 2> gpurun_out/lmm_bench_c2.err; tail -c 200 gpurun_out/lmm_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/lmm_bench_c2.json').read().strip().splitlines()[-1])
print(round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('passes_ms'), 'db1', d['config'].get('value_decode_batch_1'), d['clocks'])
PY
