#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -n 3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu_s3.json 2> gpurun_out/bench_2gpu_s3.err
echo "2gpu exit=$?"; grep -v "Warning\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_2gpu_s3.err | tail -n 5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu_s3.json').read().strip().splitlines()[-1])
print('2 GPUs: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['config']['parallelism'], 'enq', d['host_enqueue_ms_per_step'])
PY
