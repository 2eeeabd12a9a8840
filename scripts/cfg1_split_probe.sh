#!/bin/bash
# cfg1 (2^20 x 1 column) with and without the even/odd split, cold L2
export PBK_QUICK_FLUSH=1 PBK_QUICK_ITERS=20
echo "== split (default)"; python scripts/gpu_quick.py cfg1 2>&1 | grep -v "wall\|NVIDIA"
echo "== PBK_NO_SPLIT=1"; PBK_NO_SPLIT=1 python scripts/gpu_quick.py cfg1 2>&1 | grep -v "wall\|NVIDIA"
for lv in "9,10" "10,9" "8,11" "11,8" "7,6,6" "6,7,6"; do echo "== split PBK_LEVELS=$lv"; PBK_LEVELS=$lv python scripts/gpu_quick.py cfg1 2>&1 | grep -v "wall\|NVIDIA"; done
