#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
timeout 300 python scripts/gpu_stress.py 12 200 nms,fp_vote_fused 2>&1 | grep -v "Warning: CUDA warning" | tail -n 3
run() { echo "== $*"; python bench.py --steps 600 --warmup 24 --no-cpu-baseline "$@" 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1)); print('   nms', [k['ms'] for k in d['kernels'] if 'nms' in k['kernel']])"; grep -v Warning gpurun_out/err.txt | tail -n 2; }
run --inflight 12
run --inflight 12 --tune sa_sms=128
run --inflight 12 --tune sa_sms=120
run --inflight 12 --tune sa_split=2
run --inflight 12 --tune sa_sms=124 --tune sa_split=2
run --inflight 16
