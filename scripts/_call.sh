for PS in 8 4 2; do
L=$PWD/votenet_b200/libvnb_ps$PS.so; [ $PS = 8 ] && L=$PWD/votenet_b200/libvotenet_b200.so
echo "== PS=$PS"; VNB_LIB=$L timeout 200 python scripts/gpu_fps_ablate.py 2>&1 | grep "full\|setup"
VNB_LIB=$L timeout 300 python -m pytest tests/test_gpu_fps_pruned.py -x -q 2>&1 | tail -n 2
done
