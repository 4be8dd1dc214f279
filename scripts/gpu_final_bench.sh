#!/bin/bash
# End-of-round bench lines of every BASELINE config on one B200 (the driver's own command first), then the reference arm.
mkdir -p gpurun_out
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_final.json 2> gpurun_out/r02_bench_c2_final.err; tail -c 300 gpurun_out/r02_bench_c2_final.err
for c in 3 4 5; do timeout 1200 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c${c}_final.json 2> gpurun_out/r02_bench_c${c}_final.err; tail -c 300 gpurun_out/r02_bench_c${c}_final.err; done
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; tail -c 300 gpurun_out/r02_bench_reference_final.err
python - <<'PY'
import json
for c in (2,3,4,5):
    d=json.loads(open(f'gpurun_out/r02_bench_c{c}_final.json').read().strip().splitlines()[-1])
    print(c, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'db1', d['config'].get('value_decode_batch_1'), 'lat', round(d['config']['latency_ms_single_clip'],1), 'roof', round(d['roofline']['frac'],3), 'dec', d.get('roofline_decode',{}).get('ms_per_step'), 'md', d.get('mask_decoder',{}).get('ms_per_frame'), d['clocks'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
d=json.loads(open('gpurun_out/r02_bench_reference_final.json').read().strip().splitlines()[-1])
print('reference', {k: d[k] for k in ('value','unit','ms_per_step','impl') if k in d})
PY
