#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q 2>&1 | tail -15
timeout 600 python scripts/gpu_probe_sa.py 2>&1 | tee gpurun_out/probe_sa.txt
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_sa2.json 2> gpurun_out/bench_sa2.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_sa2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_sa2.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'clocks', d['clocks'])
for k in d['kernels']: print('   ', k['kernel'], k['ms'], round(k['frac'],3))
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
