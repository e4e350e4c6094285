#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sa.csv python scripts/gpu_probe_sa.py > gpurun_out/probe_sa_ncu.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa1_ws2_kernel|sa_ws2_kernel" -s 2 -c 4 -o gpurun_out/prof_sa_v2 python scripts/gpu_probe_sa.py > gpurun_out/prof_sa_v2.log 2>&1
ls -la gpurun_out/
