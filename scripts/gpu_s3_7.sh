#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --inflight 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list (no graph) exit=$?"
grep -c fps_pruned gpurun_out/launches_bench.csv; grep "ERROR" gpurun_out/launches_bench.csv | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_fwd.csv python scripts/gpu_one_forward.py 2 1 > gpurun_out/ncu_fwd.log 2>&1; echo "ncu list (one_forward) exit=$?"
grep -c fps_pruned gpurun_out/launches_fwd.csv; grep "ERROR" gpurun_out/launches_fwd.csv | head -3
