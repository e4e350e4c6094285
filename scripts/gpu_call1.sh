#!/bin/bash
# session call 1: validate head + new ball query; A/B the query variants under load; sa1 stage trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -n 2
python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_driverlike.json 2> gpurun_out/c1_bench_driverlike.err; echo "bench (driver-like) exit=$?"
python bench.py --no-cpu-baseline > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench exit=$?"
python bench.py --no-cpu-baseline --tune ball_query_variant=1 > gpurun_out/c1_bench_bq1.json 2> gpurun_out/c1_bench_bq1.err; echo "bench bq1 exit=$?"
echo "--- stress variant 2"; timeout 300 python scripts/gpu_stress.py 12 200 bq1,bq2,bq34,sa1,sa2,sa3,sa4,prop,fp_vote_fused,nms,fps1 2>&1 | tail -n 12
echo "--- stress variant 1"; VNB_TUNE=ball_query_variant=1 timeout 300 python scripts/gpu_stress.py 12 200 bq1,bq2,bq34 2>&1 | tail -n 3
echo "--- sa1 trace"; timeout 200 python scripts/gpu_trace_sa1.py 2>&1 | tail -n 34
python - <<'PY'
import json
for f in ('c1_bench_driverlike', 'c1_bench', 'c1_bench_bq1'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'lat1', d.get('latency_ms_inflight1'), 'launches', d.get('launches_per_forward'), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    if f == 'c1_bench':
        for k in d['kernels']:
            print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
