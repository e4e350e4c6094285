#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_index_ops test_gpu_dense test_gpu_engine test_gpu_forward; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -x --no-header -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 15 gpurun_out/$f.log | cut -c1-300
done
timeout 600 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('BENCH value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'lpf', d['launches_per_forward'])
    for k in d['kernels']: print('  ', k['kernel'], k['ms'], round(k['frac'],3))
except Exception as e: print('bench parse error', e)
PY
# ncu: launch list of one eager (non-graph) bench run, then full captures of the hot kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-launches exit=$?" >> gpurun_out/summary.txt
for k in fps_cluster_kernel sa_ws_kernel sa_tc_kernel grid_query_kernel linear_tc_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu-$k exit=$?" >> gpurun_out/summary.txt
done
ls -la gpurun_out | head -40
cat gpurun_out/summary.txt
