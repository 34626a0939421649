set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_radix.py -q -x 2>&1 | tail -3
for cfg in c4 c4s; do
B2Q_TRACE=1 timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-e2e --no-cpu --configs none > $O/s4_$cfg.json 2> $O/s4_$cfg.err; python - <<PY
import json
d=json.load(open("$O/s4_$cfg.json"))
print("$cfg", d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity_check",{}).get("ok"))
PY
grep "b2q" $O/s4_$cfg.err | tail -4
done
B2Q_GLOBAL_SPLIT=1 timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --no-e2e --no-cpu --configs none --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4 split=1', d['ms_per_step'], d['roofline']['kernel_ms'])"
