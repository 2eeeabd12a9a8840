#!/bin/bash
# level split x generic tile width sweep for the single-column cfg1 plan (cold L2)
export PBK_QUICK_FLUSH=1 PBK_QUICK_ITERS=20
for mg in 0 296 592 1184; do
  for lv in "" "10,10" "9,11" "11,9" "12,8" "7,7,6" "8,6,6" "6,6,8" "6,8,6"; do
    echo "== MINGRID=$mg LEVELS=${lv:-default}"
    if [ -n "$lv" ]; then export PBK_LEVELS=$lv; else unset PBK_LEVELS; fi
    PBK_GENERIC_MINGRID=$mg python scripts/gpu_quick.py cfg1 2>&1 | grep -v "wall\|NVIDIA"
  done
done
