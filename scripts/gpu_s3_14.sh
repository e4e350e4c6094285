#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_engine.py tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -n 3
timeout 300 python scripts/gpu_stress.py 12 200 fp_vote_fused,three_nn,prop_rest,bq1,bq2,bq34,fps_nested 2>&1 | grep -v "Warning: CUDA warning" | tail -n 8
python bench.py --steps 600 --warmup 24 --inflight 12 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1)); print('   fp', [(k['kernel'],k['ms']) for k in d['kernels'] if 'fp' in k['kernel']])"
