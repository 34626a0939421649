set -u
O=gpurun_out; mkdir -p $O
tools/atom_bench 30 > $O/r2_atom_bench.txt 2>&1; cat $O/r2_atom_bench.txt
for w in 0 1; do B2Q_WARP_PRIVATE=$w timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-e2e --no-cpu --configs none > $O/s2_c3_wp$w.json 2> $O/s2_c3_wp$w.err; python - <<PY
import json
d=json.load(open("$O/s2_c3_wp$w.json"))
print("c3 warp_private=$w", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d.get("parity_check",{}).get("ok"))
PY
done
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/s2_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/s2_gpu_tests.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" $O/s2_gpu_tests.log | tail -15
