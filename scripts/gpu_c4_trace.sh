#!/bin/bash
mkdir -p gpurun_out
VG_PIPE_TRACE=1 timeout 1200 python bench.py --config 4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c4_trace.json 2> gpurun_out/c4_trace.err
grep "^trace" gpurun_out/c4_trace.err | cut -c1-700
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c4_trace.json').read().strip().splitlines()[-1])
print(round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e']['passes_ms'])
PY
