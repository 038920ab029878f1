#!/bin/bash
# Counts of the Blackwell-native SASS mnemonics per kernel in the shipped library (the .so itself is git-ignored).
#   tools/sass_summary.sh > profiles/rNN_sass_summary.txt
SO=${1:-3dinfomax_b200/lib3dinfomax_b200.so}
echo "# cuobjdump -sass $SO : UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA, LDGMC / STG.MC = multimem (NVLS)"
cuobjdump -sass "$SO" 2>/dev/null | awk '
  /Function :/ { fn=$3 }
  /UTC[A-Z]*MMA/ { mma[fn]++ }
  /LDTM/ { ldtm[fn]++ }
  /UTMALDG/ { tma[fn]++ }
  /UBLKCP/ { blk[fn]++ }
  /LDGMC|MULTIMEM|\.MC/ { mc[fn]++ }
  /REDG|ATOMG/ { red[fn]++ }
  END { for (f in mma) k[f]=1; for (f in ldtm) k[f]=1; for (f in tma) k[f]=1; for (f in blk) k[f]=1; for (f in mc) k[f]=1;
        for (f in k) printf "%5d UTCMMA %4d LDTM %4d UTMALDG %3d UBLKCP %3d multimem %4d RED/ATOM  %s\n", mma[f], ldtm[f], tma[f], blk[f], mc[f], red[f], f }' | sort -k8 | c++filt 2>/dev/null | cut -c1-220
echo "# totals"
cuobjdump -sass "$SO" 2>/dev/null | grep -o "UTC[A-Z]*MMA\|LDTM\|UTMALDG\|UBLKCP\|LDGMC\|STG\.E\.MC[A-Z0-9.]*" | sort | uniq -c
