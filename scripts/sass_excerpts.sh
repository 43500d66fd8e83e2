#!/bin/sh
# Instruction evidence from the built library (B200_PROFILING.md: the SASS mnemonics that prove TMA / bulk copies / mbarriers / cluster
# barriers / 256-bit loads / system-scope flags are really in the kernels):  sh scripts/sass_excerpts.sh > profiles/r2_sass_excerpts.txt
LIB=${1:-polymer-sarw_b200/lib/libsarw_b200.so}
echo "# cuobjdump -sass $LIB | per kernel: count of each mnemonic of interest (sm_100a)"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ { f=$3; next }
{
  for (i = 1; i <= NF; i++) {
    t = $i
    if (t ~ /^(UTMALDG|UBLKCP|SYNCS|UCGABAR|LDG\.E\.ENL2\.256|LDG\.E\.256|RED\.E|ATOMG|ATOM\.E|MEMBAR|FENCE|LDGSTS|ST\.E\.[A-Z0-9.]*STRONG\.SYS|LD\.E\.[A-Z0-9.]*STRONG\.SYS|STG\.E\.[A-Z0-9.]*SYS|LDG\.E\.[A-Z0-9.]*SYS|DFMA|DMUL|DADD|MUFU\.RSQ64H|MUFU\.RCP64H)/) {
      sub(/;$/, "", t); c[f " " t]++
    }
  }
}
END { for (k in c) print c[k], k }' | sort -k2,2 -k3,3 | awk '{ if ($2 != last) { print ""; print $2; last = $2 } printf "    %6d  %s\n", $1, $3 }' | c++filt
