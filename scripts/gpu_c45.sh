#!/bin/bash
mkdir -p gpurun_out
for c in 4 5; do timeout 1200 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c${c}_final.json 2> gpurun_out/r02_bench_c${c}_final.err; tail -c 200 gpurun_out/r02_bench_c${c}_final.err; done
python - <<'PY'
import json
for c in (4,5):
    d=json.loads(open(f'gpurun_out/r02_bench_c{c}_final.json').read().strip().splitlines()[-1])
    print(c, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), d['e2e']['passes_ms'], 'db', d['config']['decode_batch'], 'db1', d['config'].get('value_decode_batch_1'))
PY
