#!/bin/bash
# Blackwell-native evidence from the built objects: counts of the SASS mnemonics that prove tcgen05 / TMEM /
# TMA bulk copies / mbarriers (B200_PROFILING.md "What proves a Blackwell-native kernel"), per object.
cd "$(dirname "$0")/.."
OUT=${1:-profiles/r2_sass_mnemonics.txt}
{
echo "cuobjdump -sass of spmf_b200/lib/*.o (sm_100a), built $(date -u +%Y-%m-%d) from $(git rev-parse --short HEAD)"
printf "%-20s %8s %8s %8s %8s %8s %8s %8s %8s %8s\n" object UTCHMMA LDTM STTM UTCBAR UBLKCP UTMALDG SYNCS HMMA ATOMS/RED
for o in spmf_b200/lib/*.o; do
  s=$(cuobjdump -sass "$o")
  c() { echo "$s" | grep -c "$1"; }
  printf "%-20s %8d %8d %8d %8d %8d %8d %8d %8d %8d\n" "$(basename $o)" $(c UTCHMMA) $(c LDTM) $(c STTM) $(c UTCBAR) $(c UBLKCP) $(c UTMALDG) $(c "SYNCS") $(c " HMMA") $(c "RED\.\|ATOMG\|ATOMS")
done
echo
echo "kernels containing UTCHMMA:"
for o in spmf_b200/lib/spmf_umma.o spmf_b200/lib/spmf_hot_tile.o; do
  cuobjdump -sass "$o" | awk '/Function :/ {f=$3} /UTCHMMA/ {n[f]++} END {for (k in n) printf "  %-110s %d\n", k, n[k]}' | c++filt | sort
done
} > "$OUT"
cat "$OUT" | cut -c1-200
