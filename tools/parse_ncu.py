import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5 and r[0].isdigit()]
cur={}
for r in rows:
    cur.setdefault(int(r[0]),{})[r[-3]]=float(r[-1].replace(',',''))
names=[l.split("  scan")[0].strip() for l in open(sys.argv[2]) if " scan " in l]
ids=sorted(cur)
for vi,name in enumerate(names):
    v=cur[ids[vi*6+5]]
    g=lambda k: v.get(k,0)
    print("%-30s dramR %.1f B/row W %.1f | L2 rd %.1f B/row | red %.2fM atom %.2fM wr %.2fM | inst/row %.2f | %.3f ms"%(name, g('dram__bytes_read.sum')/2e8, g('dram__bytes_write.sum')/2e8, g('lts__t_sectors_srcunit_tex_op_read.sum')*32/2e8, g('lts__t_sectors_op_red.sum')/1e6, g('lts__t_sectors_op_atom.sum')/1e6, g('lts__t_sectors_op_write.sum')/1e6, g('smsp__inst_executed.sum')/2e8, g('gpu__time_duration.sum')/1e6))
