#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -n 2
python bench.py > gpurun_out/bench_s3e.json 2> gpurun_out/bench_s3e.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3e.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'launches/fwd', d['launches_per_forward'])
print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k!='note'})
for k in d['kernels']: print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
