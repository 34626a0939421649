set -u
O=gpurun_out; mkdir -p $O
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | head -8
timeout 600 $T tools/multigpu_check.py > $O/s8_check_n$N.log 2>&1; tail -2 $O/s8_check_n$N.log
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > $O/s8_bench_n$N.json 2> $O/s8_bench_n$N.err; python - <<PY
import json
try:
    d=json.load(open("$O/s8_bench_n$N.json"))
    print("N=$N", d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d.get("parity_check"), {k:d["e2e"][k] for k in ("value","ms_per_step","rows_per_gpu")})
except Exception as e:
    print("bench failed", e); print(open("$O/s8_bench_n$N.err").read()[-1500:])
PY
timeout 900 $T bench.py --gpus $N --scaling strong --total-rows 8000000000 --steps 10 --warmup 3 --no-e2e --no-cpu > $O/s8_strong_n$N.json 2> $O/s8_strong_n$N.err; python - <<PY
import json
try:
    d=json.load(open("$O/s8_strong_n$N.json"))
    print("strong N=$N", d["ms_per_step"], d["value"], d.get("parity_check"))
except Exception as e:
    print("strong failed", e); print(open("$O/s8_strong_n$N.err").read()[-1500:])
PY
timeout 900 $T bench.py --gpus $N --config c4s --rows 125000000 --steps 5 --warmup 3 --no-e2e --no-cpu > $O/s8_c4s_n$N.json 2> $O/s8_c4s_n$N.err; python - <<PY
import json
try:
    d=json.load(open("$O/s8_c4s_n$N.json"))
    print("c4s N=$N", d["ms_per_step"], d["value"], d.get("parity_check"))
except Exception as e:
    print("c4s failed", e); print(open("$O/s8_c4s_n$N.err").read()[-1500:])
PY
