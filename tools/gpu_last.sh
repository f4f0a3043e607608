#!/bin/bash
# tools/gpu_last.sh <tag> — shortest useful validation: parity tests, one full ncu capture, one resident bench line
tag=${1:-last}; out=gpurun_out/$tag; mkdir -p $out
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/pytest_gpu.log
timeout 90 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_filter -c 1 -o $out/scan_c2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_full.log 2>&1
timeout 60 python bench.py --steps 20 --no-e2e --no-cpu > $out/bench_c2.json 2> $out/bench_c2.err; cut -c1-900 $out/bench_c2.json
