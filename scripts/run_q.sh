echo "== OLD"; PBK_LIBRARY=$PWD/pulsarbat_b200/libpbk_old.so python scripts/gpu_quick.py cfg2 2>&1 | grep -v "wall\|NVIDIA"
echo "== NEW"; python scripts/gpu_quick.py cfg2 2>&1 | grep -v "wall\|NVIDIA"
