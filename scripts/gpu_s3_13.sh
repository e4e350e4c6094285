#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fp_chain_kernel" -c 2 -f -o gpurun_out/full_fpchain python scripts/gpu_one_forward.py 1 1 > gpurun_out/ncu_fpchain.log 2>&1; echo "exit=$?"; tail -n 2 gpurun_out/ncu_fpchain.log
