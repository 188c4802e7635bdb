#!/bin/bash
# SASS evidence of the Blackwell-native instructions in libdrn_sm100.so (runs without a GPU): per kernel, the counts of
#   UTCHMMA / UTCHMMA.2CTA (tcgen05.mma, cta_group 1 / 2), LDTM (tcgen05.ld), UTMALDG (TMA loads), UTCBAR (tcgen05.commit),
#   SYNCS (mbarrier), HMMA (mma.sync) -- mnemonics per /opt/skills/guides/B200_PROFILING.md.
# Usage: bash scripts/sass_summary.sh > profiles/r02_sass_summary.txt
set -u
SO=${1:-drn_b200/libdrn_sm100.so}
echo "# cuobjdump -sass $SO  ($(stat -c %s "$SO") bytes, $(date -u +%FT%TZ))"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; order[++n]=fn }
  /UTCHMMA\.2CTA/ { two[fn]++ ; next }
  /UTCHMMA/ { one[fn]++ }
  /LDTM/ { ldtm[fn]++ }
  /UTMALDG/ { tma[fn]++ }
  /UTMASTG/ { tmast[fn]++ }
  /UTCBAR/ { bar[fn]++ }
  /SYNCS/ { syncs[fn]++ }
  /[^C]HMMA/ { hmma[fn]++ }
  END {
    printf "%-90s %8s %12s %6s %8s %8s %7s %6s %6s\n", "kernel", "UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA"
    for (i = 1; i <= n; i++) { f = order[i];
      if (one[f] + two[f] + ldtm[f] + tma[f] + tmast[f] + bar[f] + hmma[f] > 0)
        printf "%-90s %8d %12d %6d %8d %8d %7d %6d %6d\n", substr(f, 1, 90), one[f], two[f], ldtm[f], tma[f], tmast[f], bar[f], syncs[f], hmma[f]
      t1 += one[f]; t2 += two[f]; t3 += ldtm[f]; t4 += tma[f]; t5 += tmast[f]; t6 += bar[f]; t7 += syncs[f]; t8 += hmma[f] }
    printf "%-90s %8d %12d %6d %8d %8d %7d %6d %6d\n", "TOTAL (all kernels of the library)", t1, t2, t3, t4, t5, t6, t7, t8
  }'
