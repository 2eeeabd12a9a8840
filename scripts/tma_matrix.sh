run() { echo "== PBK_DBG=$PBK_DBG PASS=$PBK_TMA_PASS $*"; timeout 60 python scripts/tma_repro.py "$@" 2>&1 | grep -E "^ok|rror|pbk_tma" | sort | uniq -c | head -8; }
export PBK_TMA_PASS=1
for d in 1 2 3 4; do PBK_DBG=$d run 20 64 1 0 1; done
PBK_DBG=4 run 20 32 2 0 1 8,8,4
PBK_DBG=0 run 17 32 2 0 1 3,8,6
PBK_DBG=0 run 18 32 2 0 1 4,8,6
PBK_DBG=0 run 19 32 2 0 1 5,8,6
export PBK_TMA_PASS=0
PBK_DBG=0 run 18 64 2 0 1 8,10
PBK_DBG=0 run 20 64 2 0 1 8,12
PBK_DBG=0 run 20 16 2 0 1 8,12
