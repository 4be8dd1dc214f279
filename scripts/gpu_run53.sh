#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline --decode-batch 8 > gpurun_out/r53_c3_db8.json 2> gpurun_out/r53_c3_db8.err; tail -c 200 gpurun_out/r53_c3_db8.err
timeout 900 python bench.py --steps 24 --warmup 5 --no-cpu-baseline --decode-batch 8 > gpurun_out/r53_c2_db8.json 2> gpurun_out/r53_c2_db8.err; tail -c 200 gpurun_out/r53_c2_db8.err
python - <<'PY'
import json
for n in ('c3_db8','c2_db8'):
    d=json.loads(open(f'gpurun_out/r53_{n}.json').read().strip().splitlines()[-1])
    print(n, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('passes_ms'), 'db1', d['config'].get('value_decode_batch_1'))
PY
