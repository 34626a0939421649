set -u
O=gpurun_out; mkdir -p $O
tools/atom_bench 30 > $O/r2_atom_bench2.txt 2>&1; cat $O/r2_atom_bench2.txt
timeout 900 python -m pytest tests/test_gpu_radix.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --config c4s --steps 5 --warmup 3 --no-e2e --no-cpu --configs none > $O/s3_c4s.json 2> $O/s3_c4s.err; python - <<PY
import json
d=json.load(open("$O/s3_c4s.json"))
print("c4s", d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity_check",{}).get("ok"))
PY
tail -3 $O/s3_c4s.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:radix -s 6 -c 2 --csv --log-file $O/s3_c4s_ncu.csv python bench.py --config c4s --steps 1 --warmup 3 --no-cpu --no-e2e --configs none --no-parity > /dev/null 2>&1; grep -v "^==" $O/s3_c4s_ncu.csv | cut -d, -f5,13,15 | tail -4
