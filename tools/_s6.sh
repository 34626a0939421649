set -u
O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:k_agg --csv --log-file $O/s6_atom_bench_ncu.csv tools/atom_bench 30 > /dev/null 2>&1
grep -v "^==" $O/s6_atom_bench_ncu.csv | python -c "
import csv,sys,collections
d=collections.OrderedDict()
for r in csv.reader(sys.stdin):
    if len(r)>14 and r[0].isdigit(): d.setdefault((r[0],r[4][:40]),{})[r[12]]=r[14]
for k,v in d.items(): print(k, v)
" | awk 'NR%4==0'
