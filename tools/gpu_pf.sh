#!/bin/bash
# tools/gpu_pf.sh — same-box A/B of the scan kernel's L2 prefetch distance (viyadb_b200/libvgpu_pf<d>.so, -DVGPU_PF_DIST=d)
out=gpurun_out/pf; mkdir -p $out
run() {  # workload variant
  VGPU_LIB_PATH=$PWD/viyadb_b200/libvgpu_$2.so timeout 120 python bench.py --workload $1 --steps 10 --no-e2e --no-cpu --no-check --no-also --no-post 2>/dev/null | tail -1 > $out/$1_$2.json
  python -c "
import json; r=json.load(open('$out/$1_$2.json')); print('$1 $2 step %.3f kernel %.3f' % (r['ms_per_step'], r['roofline']['kernel_ms']))"
}
run c2 pf1; run c2 pf2; run c2 pf3; run c2 pf1; run c3 pf1; run c3 pf2
