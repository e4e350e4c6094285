#!/bin/bash
# session call 6: sa1 max-pool warps: alternate whole tiles (0) vs split every tile by column half (1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dense or forward" 2>&1 | tail -n 3
for tune in "sa1_e3_split=0" "sa1_e3_split=1"; do
  echo "--- stress VNB_TUNE='$tune'"; VNB_TUNE="$tune" timeout 200 python scripts/gpu_stress.py 12 300 sa1 2>&1 | tail -n 1
  echo "--- sa1 trace ($tune)"; VNB_TUNE="$tune" timeout 200 python scripts/gpu_trace_sa1.py 2>&1 | tail -n 12
done
python bench.py --no-cpu-baseline --tune sa1_e3_split=0 > gpurun_out/c6_bench0.json 2> gpurun_out/c6_bench0.err; echo "bench split=0 exit=$?"
python bench.py --no-cpu-baseline --tune sa1_e3_split=1 > gpurun_out/c6_bench1.json 2> gpurun_out/c6_bench1.err; echo "bench split=1 exit=$?"
python - <<'PY'
import json
for f in ('c6_bench0', 'c6_bench1'):
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'sa1', [k['ms'] for k in d['kernels'] if k['kernel'].startswith('sa1')])
PY
