#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_engine.py tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -n 15
timeout 300 python scripts/gpu_stress.py 12 200 fp_vote,fp_vote_fused,nms,fps_nested,bq2,bq34 2>&1 | grep -v "Warning: CUDA warning" | tail -n 8
python bench.py --steps 600 --warmup 24 --inflight 12 --no-cpu-baseline > gpurun_out/bench_s3c.json 2> gpurun_out/bench_s3c.err; echo "bench exit=$?"; grep -v Warning gpurun_out/bench_s3c.err | tail -n 3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3c.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'launches/fwd', d['launches_per_forward'])
for k in d['kernels']: print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
