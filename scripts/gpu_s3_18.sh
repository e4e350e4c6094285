#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fps_pruned.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -n 3
for args in "" "--steps 50 --warmup 5" "--inflight 16" "--inflight 10"; do
echo "== $args"
python bench.py --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), [ (k['kernel'],k['ms']) for k in d['kernels'] if k['kernel'] in ('fps_sa1',)])"
done
