#!/bin/bash
# session call 3: sa1 bulk-copy producer (no-swizzle A descriptor), sa_ws2 64-column hand-over
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dense or forward or engine or concurrent or train" 2>&1 | tail -n 12
python bench.py --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench exit=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench_driverlike.json 2> gpurun_out/c3_bench_driverlike.err; echo "bench (driver-like) exit=$?"
echo "--- stress"; timeout 300 python scripts/gpu_stress.py 12 200 sa1,sa2,sa3,sa4,prop 2>&1 | tail -n 6
echo "--- sa1 trace"; timeout 200 python scripts/gpu_trace_sa1.py 2>&1 | tail -n 18
echo "--- sa2 trace"; timeout 200 python scripts/gpu_trace_sa2.py 1 2>&1 | tail -n 14
python - <<'PY'
import json
for f in ('c3_bench_driverlike', 'c3_bench'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'lat1', d.get('latency_ms_inflight1'), 'launches', d.get('launches_per_forward'), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    if f == 'c3_bench':
        for k in d['kernels']:
            if 'sa' in k['kernel'] or 'prop' in k['kernel']: print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
