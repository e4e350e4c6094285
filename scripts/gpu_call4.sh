#!/bin/bash
# session call 4: sa1 E3 cid cache, sa_ws2 back to 4 chunks, fp_chain 256-bit I/O + TMEM residual + hoisted prologue loads, linear_tc 256-bit I/O
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 12
python bench.py --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench exit=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c4_bench_driverlike.json 2> gpurun_out/c4_bench_driverlike.err; echo "bench (driver-like) exit=$?"
echo "--- stress"; timeout 300 python scripts/gpu_stress.py 12 200 three_nn,prop_rest,bq34,sa1,sa2,sa3,sa4,prop,bq1,bq2,fps_nested,fp_vote_fused,nms,fps1 2>&1 | tail -n 15
echo "--- sa1 trace"; timeout 200 python scripts/gpu_trace_sa1.py 2>&1 | tail -n 12
echo "--- fp trace"; timeout 200 python scripts/gpu_trace_fp.py 2>&1 | tail -n 16
python - <<'PY'
import json
for f in ('c4_bench_driverlike', 'c4_bench'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'lat1', d.get('latency_ms_inflight1'), 'launches', d.get('launches_per_forward'), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    if f == 'c4_bench':
        for k in d['kernels']:
            print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
