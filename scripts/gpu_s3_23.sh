#!/bin/bash
mkdir -p gpurun_out
for i in 1; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_graph$i.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_graph$i.log 2>&1; echo "ncu (graph mode, default flags) run $i exit=$?"
grep "ERROR" gpurun_out/launches_graph$i.csv | head -3; grep -c "^\"" gpurun_out/launches_graph$i.csv; grep "^\"" gpurun_out/launches_graph$i.csv | tail -n 2 | cut -c1-250
done
