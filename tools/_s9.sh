set -u
O=gpurun_out; mkdir -p $O
tools/scatter_bench 28 > $O/r2_scatter_bench2.txt 2>&1; grep -E "copy|P= 916" $O/r2_scatter_bench2.txt
timeout 900 python -m pytest tests/test_gpu_radix.py -q -x 2>&1 | tail -3
for cfg in c4s c3; do
timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-e2e --no-cpu --configs none > $O/s9_$cfg.json 2> $O/s9_$cfg.err; python - <<PY
import json
try:
    d=json.load(open("$O/s9_$cfg.json"))
    print("$cfg", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d.get("parity_check",{}).get("ok"))
except Exception as e:
    print("$cfg failed", e); print(open("$O/s9_$cfg.err").read()[-1200:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:radix -s 6 -c 2 --csv --log-file $O/s9_c4s_ncu.csv python bench.py --config c4s --steps 1 --warmup 3 --no-cpu --no-e2e --configs none --no-parity > /dev/null 2>&1
grep -v "^==" $O/s9_c4s_ncu.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>14 and r[0].isdigit(): print('  ', r[4][:45], r[12], r[14])
"
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/s9_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/s9_gpu_tests.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" $O/s9_gpu_tests.log | tail -8
