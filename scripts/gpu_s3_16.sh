#!/bin/bash
# full GPU test suite, default bench (both arms), stress decomposition, ncu launch list + full capture of the top kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
python bench.py > gpurun_out/bench_s3d.json 2> gpurun_out/bench_s3d.err; echo "bench exit=$?"; grep -v Warning gpurun_out/bench_s3d.err | tail -n 3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3d.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline'], 'clocks', d['clocks'], 'launches/fwd', d['launches_per_forward'])
print('roofline', d['roofline'])
for k in d['kernels']: print('   %-36s %8.4f ms  frac %.3f' % (k['kernel'], k['ms'], k['frac']))
PY
timeout 400 python bench.py --impl reference --steps 16 --warmup 3 | cut -c1-700
timeout 300 python scripts/gpu_stress.py 12 200 > gpurun_out/stress_s3.txt 2>&1; grep -v Warning gpurun_out/stress_s3.txt | tail -n 24
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_s3.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fps_pruned_kernel|sa_ws2_kernel|sa1_ws2_kernel|fp_chain_kernel|grid_query_kernel|nms_clip_kernel" -c 12 -f -o gpurun_out/full_s3b python scripts/gpu_one_forward.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit=$?"
