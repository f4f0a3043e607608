#!/bin/bash
# tools/quick_bench.sh [workloads...] — short resident-only bench lines, one per workload
for w in "${@:-c2 c1 c3 c4}"; do python bench.py --workload $w --steps 5 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['config']['workload'][:3], 'rows/s %.3g'%r['value'], 'ms/step %.3f'%r['ms_per_step'], 'kernel_ms %.3f'%r['roofline']['kernel_ms'], 'frac %.3f'%r['roofline']['frac'], 'gpu_ms %.3f'%r['gpu_ms_per_step'], r['config']['groups'], r['config']['group_table'])"; done
