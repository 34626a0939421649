set -u
O=gpurun_out; mkdir -p $O
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tools/multigpu_check.py > $O/s12_check_n$N.log 2>&1; grep -E "multigpu_check ok|AssertionError|Error" $O/s12_check_n$N.log | head -5
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > $O/s12_bench_n$N.json 2> $O/s12_bench_n$N.err; python - <<PY
import json
try:
    d=json.load(open("$O/s12_bench_n$N.json"))
    print("N=$N", d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d.get("parity_check",{}).get("ok"), {k:d["e2e"][k] for k in ("value","ms_per_step","rows_per_gpu")})
except Exception as e:
    print("bench failed", e); print(open("$O/s12_bench_n$N.err").read()[-1500:])
PY
