#!/bin/bash
# tools/sass_evidence.sh — static evidence from the built library (no GPU needed): per scan-kernel instantiation the registers,
# spill bytes and the counts of the instructions that matter (UBLKPF = TMA-unit bulk L2 prefetch, LDL / STL = local-memory
# traffic, LDG.E.128 = 128-bit column loads, REDG = native reductions), then the instruction mix of the instantiation C2 runs.
lib=${1:-viyadb_b200/libvgpu.so}
echo "# cuobjdump --dump-resource-usage / -sass $lib — scan_filter_groupby_kernel<kMinCtas, kSmemTable, kPlainKeys, kFast, kConj>"
cuobjdump --dump-resource-usage $lib 2>/dev/null | awk '/Function/ {f=$2} /REG:/ {print f, $0}' | grep scan_filter | sed 's/:$//' | while read f rest; do
  f=${f%:}
  echo "$f | $(echo $rest | grep -o 'REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*\|LOCAL:[0-9]*' | tr '\n' ' ')"
done
cuobjdump -sass $lib > /tmp/_sass.txt
echo "# per instantiation: UBLKPF / LDL / STL / LDG.128 / REDG / ATOMS+ATOMG / CALL"
awk '/Function :/ {f=$3} /scan_filter_groupby/ && /Function :/ {keep=1} /Function :/ && !/scan_filter_groupby/ {keep=0}
     keep && /UBLKPF/ {a[f]++} keep && / LDL/ {b[f]++} keep && / STL/ {c[f]++} keep && /LDG[.A-Za-z0-9_]*\.128/ {d[f]++} keep && / REDG/ {e[f]++}
     keep && / ATOM/ {g[f]++} keep && / CALL/ {h[f]++} keep {seen[f]=1}
     END {for (f in seen) printf "%s  UBLKPF %d  LDL %d  STL %d  LDG.128 %d  REDG %d  ATOM %d  CALL %d\n", f, a[f], b[f], c[f], d[f], e[f], g[f], h[f]}' /tmp/_sass.txt | sort
echo "# --- static instruction mix, scan_filter_groupby_kernel<3,false,true,true,true> (no shared-memory table, plain keys, fast, conjunction: the one C2 runs) ---"
awk '/Function :/ {keep = ($3 ~ /scan_filter_groupby_kernelILi3ELb0ELb1ELb1ELb1E/)} keep' /tmp/_sass.txt | grep -E '^\s+/\*[0-9a-f]{4,}\*/\s+[@A-Z]' |
  sed -E 's/^\s*\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{split($1, a, "."); print a[1]}' | sort | uniq -c | sort -rn | head -40
rm -f /tmp/_sass.txt
