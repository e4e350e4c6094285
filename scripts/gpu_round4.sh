#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_index_ops test_gpu_dense test_gpu_engine test_gpu_forward; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -x --no-header -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 12 gpurun_out/$f.log | cut -c1-300
done
timeout 600 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/probe.log
for nf in 2 4 6; do
timeout 600 python bench.py --steps 200 --warmup 10 --inflight $nf --no-cpu-baseline > gpurun_out/bench_if$nf.json 2> gpurun_out/bench_if$nf.err; echo "bench if$nf exit=$?" >> gpurun_out/summary.txt
tail -n 3 gpurun_out/bench_if$nf.err
done
python - <<'PY'
import json
for nf in (2,4,6):
    try:
        d=json.loads(open(f'gpurun_out/bench_if{nf}.json').read().strip().splitlines()[-1])
        print('BENCH inflight',nf,'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'lpf', d['launches_per_forward'])
        if nf==4:
            for k in d['kernels']: print('  ', k['kernel'], k['ms'], round(k['frac'],3))
    except Exception as e: print('bench parse error', nf, e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu-launches exit=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_push_kernel -s 1 -c 1 -f -o gpurun_out/prof_fps_push python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 2 > gpurun_out/ncu_fps.log 2>&1; echo "ncu-fps exit=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_linear python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 2 > gpurun_out/ncu_lin.log 2>&1; echo "ncu-linear exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
