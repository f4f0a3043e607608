#!/bin/bash
# tools/gpu_round.sh <tag> — one GPU call: parity tests, the bench line, the ncu launch list of the
# bench's timed region and one full capture of the scan kernel. Everything lands in gpurun_out/<tag>/.
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
  tail -3 $out/pytest_gpu.log
fi
timeout 900 python bench.py > $out/bench_c2.json 2> $out/bench_c2.err; tail -c 3000 $out/bench_c2.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; cat $out/bench_ref.json
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_filter -c 1 \
  -o $out/scan_c2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_full.log 2>&1
for w in ${WORKLOADS:-c1 c3 c4}; do
  timeout 600 python bench.py --workload $w --steps 10 --no-e2e --no-cpu > $out/bench_$w.json 2> $out/bench_$w.err; tail -c 1500 $out/bench_$w.json
done
for tune in $VARIANTS; do   # A/B of VGPU_TUNE bits on the headline workload (and $VARIANT_WORKLOADS)
  for w in c2 $VARIANT_WORKLOADS; do
    echo "VGPU_TUNE=$tune $w"; VGPU_TUNE=$tune timeout 600 python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $out/bench_${w}_tune$tune.json | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('   rows/s %.3g'%r['value'], 'ms/step %.3f'%r['ms_per_step'], 'kernel_ms %.3f'%r['roofline']['kernel_ms'], 'frac %.3f'%r['roofline']['frac'])"
  done
done
if [ -n "$EXPLORE" ]; then timeout 600 python tools/explore.py 200000000 c2 > $out/explore_c2.txt 2>&1; cat $out/explore_c2.txt; fi
