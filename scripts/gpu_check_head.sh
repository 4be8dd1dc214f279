#!/bin/bash
# Full GPU suite + smoke + the driver's bench command on the committed state.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/head_tests.log 2>&1; echo "rc=$?" >> gpurun_out/head_tests.log; tail -3 gpurun_out/head_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/head_bench_c2.json 2> gpurun_out/head_bench_c2.err; tail -c 200 gpurun_out/head_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/head_bench_c2.json').read().strip().splitlines()[-1])
print(round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('passes_ms'), 'db1', d['config'].get('value_decode_batch_1'), 'launches', d['gpu_launches'], d['clocks'])
PY
