#!/bin/bash
# tools/gpu_ab.sh — same-box A/B of scan-kernel variants (viyadb_b200/libvgpu_<v>.so) and of the round-1 tree (_r1/)
out=gpurun_out/ab; mkdir -p $out
for w in c2 c3; do
  for v in "$@"; do
    VGPU_LIB_PATH=$PWD/viyadb_b200/libvgpu_$v.so timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu --no-check --no-also 2>/dev/null | tail -1 > $out/${w}_$v.json
    python -c "
import json; r=json.load(open('$out/${w}_$v.json')); print('$w $v step %.3f kernel %.3f' % (r['ms_per_step'], r['roofline']['kernel_ms']))"
  done
  (cd _r1 && timeout 300 python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json; r=json.loads(sys.stdin.read()); print('$w r1 step %.3f kernel %.3f' % (r['ms_per_step'], r['roofline']['kernel_ms']))")
done
