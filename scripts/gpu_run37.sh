#!/bin/bash
mkdir -p gpurun_out
VG_PIPE_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r37_bench_c2.json 2> gpurun_out/r37_bench_c2.err
grep "^trace" gpurun_out/r37_bench_c2.err | cut -c1-1500
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r37_bench_c2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('value_decode_batch_1'), d['clocks'])
PY
