#!/bin/bash
# tools/gpu_multi.sh <tag> <ngpus> — N-GPU parity check + bench lines at N GPUs (weak scaling)
tag=${1:-m}; n=${2:-2}; E2E_FLAGS=${E2E_FLAGS---no-e2e}; out=gpurun_out/$tag; mkdir -p $out
[ -n "$SKIP_CHECK" ] || timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -9 | tee $out/multi_check.log
for w in ${WORKLOADS:-c2 c3}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $n --workload $w --steps 10 --warmup 3 $E2E_FLAGS 2> $out/bench_${w}_n$n.err | grep '^{' | tail -1 | tee $out/bench_${w}_n$n.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('N=$n', r['config']['workload'][:3], 'rows/s %.3g'%r['value'], 'ms/step %.3f'%r['ms_per_step'], 'kernel_ms %.3f'%r['roofline']['kernel_ms'], 'gpu_ms %.3f'%r['gpu_ms_per_step'], 'e2e', r['e2e'] and '%.3g'%r['e2e']['value'])"
done
