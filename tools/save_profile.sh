#!/bin/bash
# tools/save_profile.sh <gpurun_out tag> <profiles name> — copy the judged evidence of a GPU call into profiles/
src=gpurun_out/$1; dst=profiles/$2
mkdir -p profiles
[ -f $src/launches_c2.csv ] && grep -v '^==' $src/launches_c2.csv | python -c "
import csv,sys,collections
rows=list(csv.DictReader(sys.stdin))
print('# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu')
print('# every launch of the timed region (2 steps), cold-cache and serialised: compare SHARES, not absolutes')
print('id,kernel,grid,block,ns')
tot=collections.OrderedDict()
for r in rows:
    k=r['Kernel Name'].split('(')[0]; ns=int(r['Metric Value'].replace(',',''))
    print(','.join([r['ID'],k,r['Grid Size'].strip('()').split(',')[0],r['Block Size'].strip('()').split(',')[0],str(ns)]))
    tot[k]=tot.get(k,[0,0]); tot[k][0]+=ns; tot[k][1]+=1
s=sum(v[0] for v in tot.values())
print('# --- share of the step by kernel ---')
for k,v in tot.items(): print('# %-32s launches %3d  total %10d ns  share %5.1f %%'%(k,v[1],v[0],100.0*v[0]/s))
" > ${dst}_launches.csv
[ -f $src/scan_c2.ncu-rep ] && { echo "# ncu --set full --clock-control none --import-source on -k regex:scan_filter -c 1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"; python tools/ncu_summary.py $src/scan_c2.ncu-rep; } > ${dst}_scan_ncu.txt
# DRAM traffic per launch of the scan kernel -> profiles/traffic.json (read by bench.py for roofline.traffic)
for w in c2 c1 c3 c4; do [ -f $src/scan_$w.ncu-rep ] && python - $src/scan_$w.ncu-rep $w "$2" <<'PY'
import csv, json, os, subprocess, sys
rep, w, name = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units, vals = rows[0], rows[1], rows[-1]
tot = 0.0
for h, u, v in zip(hdr, units, vals):
    if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
p = "profiles/traffic.json"
t = json.load(open(p)) if os.path.exists(p) else {}
t[w] = {"bytes_per_launch": tot, "source": f"profiles/{name}_scan_ncu.txt (ncu --set full, one launch)"}
json.dump(t, open(p, "w"), indent=1, sort_keys=True)
PY
done
for f in bench_c2.json bench_ref.json bench_c1.json bench_c3.json bench_c4.json explore_c2.txt; do [ -s $src/$f ] && cp $src/$f ${dst}_$f; done
ls -la profiles | tail -12
