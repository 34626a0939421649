set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_radix.py tests/test_gpu_multi.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --config c4s --steps 10 --warmup 3 --no-e2e --no-cpu --configs none > $O/s10_c4s.json 2> $O/s10_c4s.err; python - <<PY
import json
try:
    d=json.load(open("$O/s10_c4s.json"))
    print("c4s", d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity_check",{}).get("ok"))
except Exception as e:
    print("c4s failed", e); print(open("$O/s10_c4s.err").read()[-1200:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:radix -s 6 -c 2 --csv --log-file $O/s10_c4s_ncu.csv python bench.py --config c4s --steps 1 --warmup 3 --no-cpu --no-e2e --configs none --no-parity > /dev/null 2>&1
grep -v "^==" $O/s10_c4s_ncu.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>14 and r[0].isdigit(): print('  ', r[4][:45], r[12], r[14])
"
