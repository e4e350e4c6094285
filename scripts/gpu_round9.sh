#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "2gpu exit=$?"; tail -n 5 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print('2 GPUs: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['config']['parallelism'])
PY
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "1gpu exit=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print('1 GPU: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline'], 'clocks', d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 16 --warmup 3 | cut -c1-600
