#!/bin/bash
# tools/sass_lines.sh <mangled-substring> — static SASS of one kernel of libvgpu.so by source line: spills (LDL/STL) and instruction counts
rm -rf /tmp/cub && mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all /root/repo/viyadb_b200/libvgpu.so > /dev/null 2>&1 && nvdisasm --print-line-info vgpu.sm_100a.cubin > /tmp/all_lines.txt 2>/dev/null)
python3 - "$1" <<'PY'
import re, sys
pat = sys.argv[1]
cur=None; counts={}; tot={}; infn=False
for l in open('/tmp/all_lines.txt'):
    if '.text.' in l and ('.section' in l or l.startswith('.text.')):
        infn = pat in l
        continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.search(r'^\s+/\*[0-9a-f]+\*/',l):
        tot[cur]=tot.get(cur,0)+1
        if 'LDL' in l or 'STL' in l: counts[cur]=counts.get(cur,0)+1
print('static instructions', sum(tot.values()), 'LDL/STL', sum(counts.values()))
for k,v in sorted(counts.items(), key=lambda kv:-kv[1])[:30]: print('  spill', k, v)
PY
