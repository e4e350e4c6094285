#!/bin/bash
# compute-sanitizer over the kernels of one full forward (run on a B200 box):  bash scripts/gpu_sanitize.sh [tag]
# memcheck at the bench shape; racecheck / synccheck / initcheck on the small shape (same kernels, two forwards in
# flight on two streams).  synccheck cannot instrument the tcgen05 kernels (see profiles/README.md): they are excluded
# from the main synccheck pass and run alone in a second pass whose outcome is recorded as is.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
TC='regex=sa1_ws2_kernel|sa_ws2_kernel|fp_chain_kernel|linear_tc_kernel'
summ() { grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok: kept|Error|error|hazard" "$1" | sort | uniq -c | head -n 20; }
run() {  # name, timeout, tool args..., -- python args
  name=$1; shift; to=$1; shift
  ( time timeout $to compute-sanitizer "$@" ) > $OUT/san_${TAG}_$name.txt 2>&1
  echo "== $name exit=$?"; summ $OUT/san_${TAG}_$name.txt
}
run memcheck_small 300 --tool memcheck python scripts/gpu_one_forward.py 2 2 small
run memcheck_bench 400 --tool memcheck python scripts/gpu_one_forward.py 1 1 bench
run racecheck_small 400 --tool racecheck --racecheck-report all python scripts/gpu_one_forward.py 2 2 small
run initcheck_small 300 --tool initcheck python scripts/gpu_one_forward.py 1 1 small
run synccheck_small 300 --tool synccheck --kernel-name-exclude "$TC" python scripts/gpu_one_forward.py 2 2 small
run synccheck_tcgen05_only 200 --tool synccheck --kernel-name "$TC" python scripts/gpu_one_forward.py 1 1 small
