for fam in 16 8; do
  echo "== FAMILY $fam"
  PBK_FAMILY=$fam python scripts/gpu_quick.py cfg5_shard 2>&1 | grep -v "^NVIDIA\|wall"
done
echo "== FAMILY 16 LEVELS 9,9,8"; PBK_LEVELS=9,9,8 python scripts/gpu_quick.py cfg5_shard 2>&1 | grep -v "^NVIDIA\|wall"
echo "== FAMILY 16 LEVELS 8,9,9"; PBK_LEVELS=8,9,9 python scripts/gpu_quick.py cfg5_shard 2>&1 | grep -v "^NVIDIA\|wall"
