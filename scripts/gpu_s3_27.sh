#!/bin/bash
# final profile refresh: ncu launch list of the bench command (graph mode, default flags) + --set full of the top kernels + stress table
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --inflight 2 > gpurun_out/ncu_final.log 2>&1; echo "ncu list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fps_pruned_kernel|sa_ws2_kernel|sa1_ws2_kernel|fp_chain_kernel|grid_query_kernel|nms_clip_kernel|nms_pairs_kernel|linear_tc_kernel" -c 16 -f -o gpurun_out/full_final python scripts/gpu_one_forward.py 1 1 > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full exit=$?"
timeout 300 python scripts/gpu_stress.py 12 200 > gpurun_out/stress_final.txt 2>&1; grep -v Warning gpurun_out/stress_final.txt | tail -n 22
