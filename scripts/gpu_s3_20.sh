#!/bin/bash
mkdir -p gpurun_out
for mn in 4096 1024 512; do
echo "== bq_grid_min_n=$mn"
VNB_TUNE=bq_grid_min_n=$mn timeout 300 python scripts/gpu_stress.py 12 200 bq2,bq34 2>&1 | grep -v "Warning: CUDA warning" | tail -n 2
python bench.py --no-cpu-baseline --tune bq_grid_min_n=$mn 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), [ (k['kernel'],k['ms']) for k in d['kernels'] if 'ball' in k['kernel']])"
done
timeout 600 python -m pytest tests/test_gpu_index_ops.py -m gpu -x -q -k "ball or grid or query" 2>&1 | tail -n 2
