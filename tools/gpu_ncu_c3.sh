out=gpurun_out/r1g; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_filter -c 1 -o $out/scan_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_filter -c 1 -o $out/scan_c1 python bench.py --workload c1 --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_c1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pairs_dedupe -c 1 -o $out/dedupe_c2 python bench.py --workload c2 --rows 200000000 --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_dd.log 2>&1
