#!/bin/bash
# Round-2 evidence session on ONE B200 (run through gpurun): GPU tests, the default bench line, the ncu launch list of the
# bench command and one `ncu --set full` capture of the dominant kernel per configuration.  Outputs under gpurun_out/
# (the reports are summarised ON the box with tools/ncu_summary.py; only the c2 / c4 / c4s reports travel back — 64 MiB cap).
set -u
O=gpurun_out
rm -rf $O; mkdir -p $O
B="--no-cpu --no-e2e --configs none --no-parity"
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -q -rf > $O/r2_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2_gpu_tests.log
  grep -E "^(FAILED|ERROR)|passed|failed|rc=" $O/r2_gpu_tests.log | tail -15
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err; tail -c 300 $O/r2_bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2_raw.csv python bench.py --steps 2 --warmup 3 $B > /dev/null 2> $O/r2_launches.err
for cfg in c3 c4 c4s; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv --log-file $O/r2_launches_${cfg}_raw.csv python bench.py --config $cfg --steps 1 --warmup 3 $B > /dev/null 2>> $O/r2_launches.err
done
bpr() { case $1 in c2) echo 20;; c2all) echo 36;; c3) echo 20;; *) echo 16;; esac; }
for cfg in c2 c2all c3 c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2q_k_scan -s 3 -c 1 -f -o $O/r2_scan_${cfg}_full python bench.py --config $cfg --steps 1 --warmup 3 $B > /dev/null 2> $O/r2_ncu_${cfg}.err
  python tools/ncu_summary.py $O/r2_scan_${cfg}_full.ncu-rep 1000000000 $(bpr $cfg) > $O/r2_scan_${cfg}_full_ncu.txt 2>> $O/r2_ncu_${cfg}.err
  ncu -i $O/r2_scan_${cfg}_full.ncu-rep --page raw --csv > $O/r2_scan_${cfg}_full_raw.csv 2>> $O/r2_ncu_${cfg}.err
done
rm -f $O/r2_scan_c2all_full.ncu-rep $O/r2_scan_c3_full.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix -s 6 -c 2 -f -o $O/r2_radix_c4s_full python bench.py --config c4s --steps 1 --warmup 3 $B > /dev/null 2> $O/r2_ncu_c4s.err
python tools/ncu_summary.py $O/r2_radix_c4s_full.ncu-rep 1000000000 16 > $O/r2_radix_c4s_full_ncu.txt 2>> $O/r2_ncu_c4s.err
ncu -i $O/r2_radix_c4s_full.ncu-rep --page raw --csv > $O/r2_radix_c4s_full_raw.csv 2>> $O/r2_ncu_c4s.err
du -sh $O; ls -la $O | tail -30
