#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_forward.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -n 3
python scripts/gpu_trace_sa1.py 2>&1 | grep -v Warning | tail -n 7
timeout 300 python scripts/gpu_stress.py 12 200 sa1 2>&1 | grep -v "Warning: CUDA warning" | tail -n 1
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), [ (k['kernel'],k['ms'],round(k['frac'],3)) for k in d['kernels'] if k['kernel'] in ('sa1_group_mlp_max',)])"
