#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_sam2_gpu.py -x -q > gpurun_out/r45_check.log 2>&1; echo "rc=$?" >> gpurun_out/r45_check.log
tail -3 gpurun_out/r45_check.log
timeout 300 python scripts/gpu_attn_tc_check.py 2>&1 | grep "bench" | grep "legacy=False" | head -2
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r45_bench_c2.json 2> gpurun_out/r45_bench_c2.err; tail -c 300 gpurun_out/r45_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r45_bench_c2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('value_decode_batch_1'), d['config']['latency_ms_single_clip'], d['roofline']['frac'], d['clocks'])
PY
