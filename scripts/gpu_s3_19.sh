#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fps_pruned.py tests/test_gpu_dense.py -m gpu -x -q 2>&1 | tail -n 3
timeout 300 python scripts/gpu_stress.py 12 200 sa2,sa3,sa4,prop,fps1 2>&1 | grep -v "Warning: CUDA warning" | tail -n 6
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), [ (k['kernel'],k['ms']) for k in d['kernels'] if k['kernel'] in ('fps_sa1','sa2_group_mlp_max','sa3_group_mlp_max','sa4_group_mlp_max')])"
