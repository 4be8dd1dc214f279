#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_n2_final.json 2> gpurun_out/r02_bench_c2_n2_final.err; tail -c 400 gpurun_out/r02_bench_c2_n2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_n2_final.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('passes_ms'), d['clocks'])
PY
