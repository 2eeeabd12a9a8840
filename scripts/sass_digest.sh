#!/bin/bash
# SASS evidence (VERDICT r1 missing #7): instruction mnemonics of every compute kernel in
# libpbk.so -- packed FP32 (FFMA2/FADD2/FMUL2: Blackwell), TMA (UTMALDG), mbarrier (SYNCS),
# no tensor-core instructions on this bandwidth-bound path.  Run from the repo root after a build.
set -e
SO=pulsarbat_b200/libpbk.so
OUT=${1:-profiles/r02_sass_digest.txt}
{
  echo "cuobjdump -sass $SO  ($(date -u +%F), $(nvcc --version | tail -1))"
  echo
  echo "== whole library: instruction counts of interest"
  cuobjdump -sass $SO | awk '{print $2}' | grep -E '^(FFMA2|FADD2|FMUL2|FFMA|FADD|FMUL|DFMA|DADD|DMUL|UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|LDGSTS[.A-Z0-9]*|LDG[.A-Z0-9]*|STG[.A-Z0-9]*|LDS[.A-Z0-9]*|STS[.A-Z0-9]*|SHFL[.A-Z0-9]*|BAR[.A-Z0-9]*|MUFU[.A-Z0-9]*|RED[.A-Z0-9]*|ATOM[.A-Z0-9]*|HMMA[.A-Z0-9]*|UTC[A-Z0-9.]*MMA[.A-Z0-9]*|LDTM|STTM)$' \
    | sed -E 's/^(LDG|STG|LDS|STS|RED|ATOM[GS]?|SHFL|BAR|MUFU|SYNCS|UTMALDG|LDGSTS)\..*/\1.*/' | sort | uniq -c | sort -rn
  echo
  echo "== per kernel (mangled template arguments kept short): FFMA2+FADD2+FMUL2 / scalar FP32 / FP64 / UTMALDG / SYNCS / BAR"
  cuobjdump -sass $SO | awk '
    /Function : / { if (name != "") printf "%-110s %6d %6d %6d %4d %4d %4d\n", name, p2, p1, d, t, y, b;
                    name=$3; gsub(/_ZN3pbk/, "", name); name=substr(name, 1, 110); p2=p1=d=t=y=b=0 }
    { m=$2 }
    m ~ /^(FFMA2|FADD2|FMUL2)$/ { p2++ }
    m ~ /^(FFMA|FADD|FMUL)(\.|$)/ { p1++ }
    m ~ /^(DFMA|DADD|DMUL)(\.|$)/ { d++ }
    m ~ /^UTMALDG/ { t++ }
    m ~ /^SYNCS/ { y++ }
    m ~ /^BAR/ { b++ }
    END { if (name != "") printf "%-110s %6d %6d %6d %4d %4d %4d\n", name, p2, p1, d, t, y, b }' \
    | grep -E "pass_kernel|fold|detect|f64_|blue|chirp|stokes|downsample" | sort
} > "$OUT"
echo "wrote $OUT ($(wc -l < "$OUT") lines)"
