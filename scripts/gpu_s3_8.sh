#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
timeout 300 python scripts/gpu_stress.py 12 300 2>&1 | grep -v "Warning: CUDA warning" | tail -n 16
python bench.py --steps 400 --warmup 16 --no-cpu-baseline > gpurun_out/bench_s3b.json 2> gpurun_out/bench_s3b.err; echo "bench exit=$?"; tail -n 3 gpurun_out/bench_s3b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3b.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'clocks', d['clocks'], 'enq', d['host_enqueue_ms_per_step'])
for k in d['kernels']: print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
