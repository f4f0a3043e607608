#!/bin/bash
# tools/gpu_n8_debug.sh — the C2 multi-GPU check at N GPUs under NCCL variants (which exchange loses updates?)
n=${1:-8}; out=gpurun_out/r2n; mkdir -p $out
run() { # label, env...
  label=$1; shift
  env VGPU_CHECK_ONLY=c2 "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py > $out/check_$label.log 2>&1
  echo "== $label: $(grep -c 'identical on' $out/check_$label.log) ok lines; $(grep '^\[rank 0\]' $out/check_$label.log | head -3)"
}
run default
run ungroup VGPU_NCCL_UNGROUP=1
run ring NCCL_ALGO=Ring
run nonvls NCCL_NVLS_ENABLE=0
