#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sam2_gpu.py tests/test_e2e_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/sam_tests.log 2>&1; echo "rc=$?" >> gpurun_out/sam_tests.log; tail -4 gpurun_out/sam_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/sam_bench_c2.json 2> gpurun_out/sam_bench_c2.err; tail -c 200 gpurun_out/sam_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/sam_bench_c2.json').read().strip().splitlines()[-1])
print(round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('passes_ms'), d['clocks'])
PY
