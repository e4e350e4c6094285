#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_dense test_gpu_forward test_gpu_engine; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 600 -x --no-header -p no:cacheprovider -s > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  grep -E "rel err|passed|failed|Error|error" gpurun_out/$f.log | tail -n 40
done
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
tail -c 6000 gpurun_out/bench.json; tail -n 20 gpurun_out/bench.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench-nograph exit=$?" >> gpurun_out/summary.txt
python -c "
import json
for f in ('gpurun_out/bench.json','gpurun_out/bench_nograph.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e'], d['launches_per_forward'])
    except Exception as e: print(f, 'ERR', e)
"
cat gpurun_out/summary.txt
