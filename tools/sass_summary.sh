#!/bin/bash
# SASS evidence per object of libptrec_b200: counts of the tensor-core / TMEM / TMA / bulk-copy / memory mnemonics
# (B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld,
# UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, HMMA = legacy mma.sync).   tools/sass_summary.sh > profiles/r2_sass_summary.txt
cd "$(dirname "$0")/../pytorchrec_b200/csrc/build" || exit 1
printf "%-16s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" object UTCHMMA UTMALDG UTMASTG LDTM UTCBAR UBLKCP HMMA LDG.128 STG.128
for o in *.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -c "$1"; }
  printf "%-16s %8d %8d %8d %8d %8d %8d %8d %8d %8d\n" "$o" "$(c UTCHMMA)" "$(c UTMALDG)" "$(c UTMASTG)" "$(c LDTM)" \
    "$(c UTCBAR)" "$(c UBLKCP)" "$(c 'HMMA')" "$(c 'LDG.E.128')" "$(c 'STG.E.128')"
done
echo
echo "# variants seen (first few distinct mnemonics per family)"
for o in tc_linear.o dcn_cross.o din_attn_tc.o gather_pool.o; do
  [ -f "$o" ] || continue
  echo "## $o"
  cuobjdump -sass "$o" 2>/dev/null | grep -oE "(UTCHMMA|UTMALDG|UTMASTG|LDTM|UTCBAR|UBLKCP|UTCCP)[A-Za-z0-9_.]*" | sort | uniq -c | sort -rn | head -12
done
