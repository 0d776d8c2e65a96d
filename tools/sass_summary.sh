#!/bin/bash
# Opcode evidence for the Blackwell-native path: counts of the SASS mnemonics of tcgen05 / TMEM / TMA in libhd_b200.so.
# usage: tools/sass_summary.sh > profiles/r02_sass_opcodes.txt
SO=${1:-human_dynamics_b200/libhd_b200.so}
echo "# cuobjdump -sass $SO  (sm_100a)  -- $(date -u +%Y-%m-%dT%H:%MZ), source $(git rev-parse --short HEAD 2>/dev/null)"
TMP=$(mktemp)
cuobjdump -sass "$SO" > "$TMP"
for op in UTCHMMA UTCQMMA LDTM STTM UTMALDG UTMASTG UTMAPF UBLKCP UTCBAR SYNCS LDGSTS ARRIVES.LDGSTSBAR USETMAXREG HMMA HGMMA; do
  printf "%-20s %6d\n" "$op" "$(grep -c "[[:space:]]$op" "$TMP")"
done
echo
echo "# per kernel (tcgen05 MMA / TMEM load / TMA load / TMA store / cp.async)"
awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /LDTM/ {l[name]++} /UTMALDG/ {t[name]++} /UTMASTG/ {s[name]++} /LDGSTS/ {c[name]++}
     END {for (k in m) printf "%5d %5d %5d %5d %5d  %s\n", m[k], l[k], t[k], s[k], c[k], k}' "$TMP" | sort -k6 | c++filt | cut -c1-200
rm -f "$TMP"
