python scripts/gpu_quick.py cfg2 cfg2_c64 cfg3_1gpu mid cfg1 n16 2>&1 | grep -v "wall"
