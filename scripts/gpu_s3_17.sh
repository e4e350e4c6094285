#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q -k "fp" 2>&1 | tail -n 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu_s3.json 2> gpurun_out/bench_2gpu_s3.err
echo "2gpu exit=$?"; grep -v "Warning\|^$" gpurun_out/bench_2gpu_s3.err | tail -n 5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu_s3.json').read().strip().splitlines()[-1])
print('2 GPUs: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['config']['parallelism'], 'launches', d['gpu_launches'])
PY
python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_s3.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu_s3.json').read().strip().splitlines()[-1])
print('1 GPU: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), [ (k['kernel'],k['ms']) for k in d['kernels'] if 'fused' in k['kernel']])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 8 --warmup 3 | cut -c1-300
