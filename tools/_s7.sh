set -u
O=gpurun_out; mkdir -p $O
for sp in 0 1; do
B2Q_GLOBAL_SPLIT=$sp timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-e2e --no-cpu --configs none > $O/s7_c4_$sp.json 2>/dev/null; python - <<PY
import json
d=json.load(open("$O/s7_c4_$sp.json"))
print("c4 split=$sp", d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity_check",{}).get("ok"))
PY
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
B2Q_GLOBAL_SPLIT=0 timeout 600 ncu --metrics $M --clock-control none -k regex:b2q_k_scan -s 3 -c 1 --csv --log-file $O/s7_c4_ncu.csv python bench.py --config c4 --steps 1 --warmup 3 --no-cpu --no-e2e --configs none --no-parity > /dev/null 2>&1
grep -v "^==" $O/s7_c4_ncu.csv | cut -d, -f13,15 | tail -3
B2Q_GLOBAL_SPLIT=0 timeout 600 python bench.py --config c4top --steps 5 --warmup 3 --no-e2e --no-cpu --configs none --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4top split=0', d['ms_per_step'], d['roofline']['kernel_ms'])"
