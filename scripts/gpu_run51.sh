#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 3; do
VG_PIPE_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r51_$r.json 2> gpurun_out/r51_$r.err
grep "^trace e2e" gpurun_out/r51_$r.err | cut -c1-900
python - <<PY
import json
d=json.loads(open('gpurun_out/r51_$r.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['with_fp32_sam_inputs']['value'], d['config'].get('value_decode_batch_1'))
PY
done
