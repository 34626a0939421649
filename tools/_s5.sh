set -u
O=gpurun_out; mkdir -p $O
B="--no-cpu --no-e2e --configs none --no-parity"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed
for sp in 0 1; do
B2Q_GLOBAL_SPLIT=$sp timeout 600 ncu --metrics $M --clock-control none -k regex:b2q_k_scan -s 3 -c 1 --csv --log-file $O/s5_c4_split$sp.csv python bench.py --config c4 --steps 1 --warmup 3 $B > /dev/null 2>&1
echo "== split=$sp"; grep -v "^==" $O/s5_c4_split$sp.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>14 and r[0].isdigit(): print('  ', r[12], r[14], r[13])
"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_aggregate -s 3 -c 1 -f -o $O/s5_pass2 python bench.py --config c4s --steps 1 --warmup 3 $B > /dev/null 2>&1
ls -la $O/s5_pass2.ncu-rep
