#!/bin/bash
# Counts of the tcgen05 / TMA / TMEM / mbarrier SASS mnemonics per kernel of one object file:
#   scripts/sass_summary.sh build/obj/attention_bwd.o
cuobjdump -sass "$1" | grep -E "Function|UTC|UTMA|LDTM|STTM|SYNCS|MUFU|LDGSTS" | awk '
/Function/ {fn=$3; sub(/^_ZN2vs[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]+[0-9][0-9]/,"",fn); next}
{for(i=1;i<=NF;i++) if ($i ~ /^(UTC|UTMA|LDTM|STTM|SYNCS|MUFU|LDGSTS)/) {m=$i; sub(/;$/,"",m); c[fn" "m]++}}
END {for (k in c) print k, c[k]}' | sort
