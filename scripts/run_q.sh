for mb in 0 16 32 64; do
  for lv in 6,8,8 7,7,8 8,7,7 8,8,6; do
    echo "== CHUNK_MB=$mb LEVELS=$lv"
    PBK_L2_CHUNK_MB=$mb PBK_LEVELS=$lv python scripts/gpu_quick.py cfg2 2>&1 | grep -v "^NVIDIA\|wall"
  done
done
