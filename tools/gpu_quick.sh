#!/bin/bash
# tools/gpu_quick.sh <tag> — parity tests + short resident-only bench lines (env VARS="A=1;B=2" adds variants)
tag=${1:-q}; out=gpurun_out/$tag; mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/pytest_gpu.log; fi
run() { # label, workload
  python bench.py --workload $2 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $out/bench_$2_$1.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('$1', r['config']['workload'][:3], 'rows/s %.3g'%r['value'], 'ms/step %.3f'%r['ms_per_step'], 'kernel_ms %.3f'%r['roofline']['kernel_ms'], 'frac %.3f'%r['roofline']['frac'])"
}
for w in ${WORKLOADS:-c2 c1 c3}; do run default $w; done
IFS=';' read -ra VS <<< "$VARS"
for v in "${VS[@]}"; do for w in ${VARIANT_WORKLOADS:-c2}; do ( export $v; run "$v" $w ); done; done
if [ -n "$NCU_WORKLOAD" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_filter -c 1 -o $out/scan_$NCU_WORKLOAD python bench.py --workload $NCU_WORKLOAD --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_$NCU_WORKLOAD.log 2>&1
fi
