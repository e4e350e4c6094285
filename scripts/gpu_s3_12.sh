#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_index_ops.py tests/test_gpu_engine.py tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -n 3
timeout 300 python scripts/gpu_stress.py 12 200 nms 2>&1 | grep -v "Warning: CUDA warning" | tail -n 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms_|rank_emit" -c 8 --csv --log-file gpurun_out/nms_list.csv python scripts/gpu_one_forward.py 1 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/nms_list.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print(r['Kernel Name'][:22], r['Grid Size'], r['Metric Name'], r['Metric Value'])"
python bench.py --steps 600 --warmup 24 --inflight 12 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1)); print('   nms', [k['ms'] for k in d['kernels'] if 'nms' in k['kernel']])"
