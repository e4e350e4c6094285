#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fps_pruned.py tests/test_gpu_index_ops.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -n 12
python bench.py --steps 600 --warmup 24 --inflight 12 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1)); print('   fps', [(k['kernel'],k['ms']) for k in d['kernels'] if 'fps' in k['kernel']])"
