for T in peer nccl; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline --transport $T > gpurun_out/r2_bench_8gpu_$T.json 2> gpurun_out/r2_bench_8gpu_$T.err; echo "bench $T exit=$?"; grep -v "Warn\|warn\|^\*\|OMP" gpurun_out/r2_bench_8gpu_$T.err | tail -3
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_8gpu_$T.json').read().strip().splitlines()[-1])
print('$T 8gpu: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), d['config']['parallelism'][-30:], d['host_enqueue_ms_per_step'])
PY
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_8gpu_driver.json 2> gpurun_out/r2_bench_8gpu_driver.err; echo "bench driver exit=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_8gpu_driver.json').read().strip().splitlines()[-1])
print('driver-like 8gpu: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1))
PY
