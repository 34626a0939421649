set -u
O=gpurun_out; mkdir -p $O
B="--no-cpu --no-e2e --configs none --no-parity"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_c2_final.json 2> $O/r2_bench_c2_final.err; tail -c 250 $O/r2_bench_c2_final.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv --log-file $O/r2_launches_c4s_final_raw.csv python bench.py --config c4s --steps 1 --warmup 3 $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix -s 6 -c 2 -f -o $O/r2_radix_c4s_final python bench.py --config c4s --steps 1 --warmup 3 $B > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_radix_c4s_final.ncu-rep 1000000000 16 > $O/r2_radix_c4s_final_ncu.txt 2>/dev/null
ncu -i $O/r2_radix_c4s_final.ncu-rep --page raw --csv > $O/r2_radix_c4s_final_raw.csv 2>/dev/null
rm -f $O/r2_radix_c4s_final.ncu-rep
grep -E "launch|gpu__time_duration|thread-instructions|traffic" $O/r2_radix_c4s_final_ncu.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2>/dev/null; head -c 400 $O/r2_bench_reference.json
