#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r34_bench_c2.json 2> gpurun_out/r34_bench_c2.err; tail -c 600 gpurun_out/r34_bench_c2.err
python - <<'PY'
import json
for c in (2,):
    d=json.loads(open(f'gpurun_out/r34_bench_c{c}.json').read().strip().splitlines()[-1])
    print(c, d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('value_decode_batch_1'), d['config']['latency_ms_single_clip'], d['gpu_launches'])
PY
for c in 4 5; do timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/r34_bench_c$c.json 2> gpurun_out/r34_bench_c$c.err; tail -c 300 gpurun_out/r34_bench_c$c.err; done
python - <<'PY'
import json
for c in (4,5):
    d=json.loads(open(f'gpurun_out/r34_bench_c{c}.json').read().strip().splitlines()[-1])
    print(c, d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('value_decode_batch_1'), d['config']['latency_ms_single_clip'], d['gpu_launches'])
PY
