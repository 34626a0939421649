#!/usr/bin/env python
"""bench.py — rows/s and achieved HBM GB/s of the scan -> filter -> group-by/aggregate path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--rows R] [--config c2|c2all|c3|c4]

A "step" is one pass of the hot path over one batch of synthetic input (BASELINE.json configs[1] at N = 1):
    1e9 rows, c0/c1 int64 ~U[0,1e6), g int32 ~U[0,1e4):  SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g
    30 fragments of 32 Mi rows (reference default fragment size, Fragmenter/FragmentDefaultValues.h:19).

`value`     whole-job rows/s with the columns already resident in HBM (timed region = the full C-ABI call:
            table init + scan + merge of CTA tables + materialise + D2H of the result buffer).
`roofline`  algorithmic bytes (20 B/row) / mean scan-kernel time (CUDA events on the launching stream, measured
            inside libb2q around the kernel) against MEASURED_PEAKS.json's HBM copy bandwidth.
`e2e`       the same query through the C ABI with HOST (pinned) column buffers: H2D copies of every referenced column
            inside the timed region, result read back to the host.
`cpu_baseline`  the oracle (CPU restatement of the reference's executor) on a bounded sample, all host threads.

N > 1 (torchrun): fragments are sharded over ranks (weak scaling: every rank scans `rows` rows of its own
fragments, fragment ids are global), partial aggregate tables are merged with NCCL all-reduce (one per dense array),
time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 0x5EED
FRAG_ROWS = 1 << 25  # 32 Mi rows

CONFIGS = {
    # name: (columns [(name, type, lo, span)], sql template, algorithmic bytes/row, workload string)
    "c2": ([("c0", "i64", 0, 10**6), ("c1", "i64", 0, 10**6), ("g", "i32", 0, 10**4)],
           "SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;", 20,
           "configs[1]: filter c0<k (50%) + GROUP BY int32 g (1e4 groups) SUM(int64)/COUNT, 3 of the 4 int64 columns unused by the minimal query are not read"),
    "c2all": ([("c0", "i64", 0, 10**6), ("c1", "i64", 0, 10**6), ("c2", "i64", 0, 10**6), ("c3", "i64", 0, 10**6), ("g", "i32", 0, 10**4)],
              "SELECT g, SUM(c1), SUM(c2), SUM(c3), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;", 36,
              "configs[1] all-columns variant: 4 x int64 + int32 key, SUM x3 + COUNT"),
    "c3": ([("f", "i64", 0, 10**6), ("g", "i32", 0, 256), ("v", "f64", 0, 1)],
           "SELECT g, AVG(v) FROM t WHERE f < 500000 GROUP BY g;", 20, "configs[2]: 256 groups AVG(double)"),
    "c4": ([("key", "i64", 0, 10**7), ("v", "i64", 0, 10**6)],
           "SELECT key, SUM(v) FROM t GROUP BY key;", 16, "configs[3]: 1e7 dense int64 keys SUM (HBM/L2 table)"),
    # sparse keys (key = U[0,1e7) * 900000000007): the range is too wide for a perfect hash => baseline hash,
    # open addressing in HBM, entry_count = NDV * 1.5 like RelAlgExecutor's estimator path
    "c4s": ([("key", "i64", 0, 10**7, 900_000_000_007), ("v", "i64", 0, 10**6)],
            "SELECT key, SUM(v) FROM t GROUP BY key;", 16, "configs[3] sparse variant: 1e7 sparse int64 keys SUM (global-memory hash table, MurmurHash3 + CAS)"),
}
CONFIGS["c2enc"] = (
    # the same table declared with the reference's fixed-width encodings (BIGINT ENCODING FIXED(32), INT ENCODING
    # FIXED(16)): 10 B/row instead of 20 — SURVEY.md §8f-2 "lets the kernels read real HeavyDB chunks"
    [("c0", "i64", 0, 10**6, 1, 4), ("c1", "i64", 0, 10**6, 1, 4), ("g", "i32", 0, 10**4, 1, 2)],
    CONFIGS["c2"][1], 10, "configs[1] with ENCODING FIXED chunks: c0,c1 BIGINT FIXED(32), g INT FIXED(16)")
# the same aggregations ending in ORDER BY ... LIMIT (top-k on the device: only the kept rows are copied back)
CONFIGS["c2top"] = (CONFIGS["c2"][0], "SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g ORDER BY 2 DESC, 1 LIMIT 10;", 20,
                    "configs[1] + ORDER BY SUM DESC LIMIT 10 (device compaction + radix sort + gather)")
CONFIGS["c4top"] = (CONFIGS["c4"][0], "SELECT key, SUM(v) FROM t GROUP BY key ORDER BY 2 DESC, 1 LIMIT 10;", 16,
                    "configs[3] + ORDER BY SUM DESC LIMIT 10 over 1e7 groups (device compaction + radix sort + gather)")
# star join — the fact table probes a 1e5-row dimension through a one-to-one perfect join table and
# groups by a dimension attribute (the dimension is generated on the host: id = row, attr = splitmix64(row) % 1000)
CONFIGS["c2join"] = ([("c0", "i64", 0, 10**6), ("c1", "i64", 0, 10**6), ("fk", "i32", 0, 10**5)],
                     "SELECT d.attr, SUM(t.c1), COUNT(*) FROM t JOIN d ON t.fk = d.id WHERE t.c0 < 500000 GROUP BY d.attr;", 20,
                     "configs[1] shape through a star join: filter c0<k (50%), INNER JOIN dim(1e5 rows) ON fk = id, GROUP BY dim.attr (1000 groups), SUM/COUNT")
CONFIGS["c2joins"] = ([("c0", "i64", 0, 10**6), ("c1", "i64", 0, 10**6), ("fk", "i32", 0, 10**4)],
                      CONFIGS["c2join"][1], 20,
                      "star join with a 1e4-row dimension: the (packed) join table is TMA-staged into shared memory")
# the cardinality-estimation query that precedes c4s in the reference's flow (CardinalityEstimationRequired ->
# RelAlgExecutor::getNDVEstimation): NDVEstimator over the sparse key, 8 B/row
CONFIGS["c4sndv"] = ([("key", "i64", 0, 10**7, 900_000_000_007)], "ESTIMATOR NDV(key)", 8,
                     "NDV estimator query over 1e7 sparse int64 keys (linear_probabilistic_count into a 1 MiB bitmap)")
ENTRY_GUESS = {"c4s": 15_000_000}


def make_unit(cfg, sql, table, names):
    """The execution unit of a config: parsed SQL, or the estimator unit (no SQL form: RelAlgExecutor synthesises it)."""
    from heavydb_b200 import abi, sqlmini
    if sql.startswith("ESTIMATOR"):
        b = abi.UnitBuilder(table)
        b.estimator([names.index("key")])
        return b.build()
    return sqlmini.parse(sql, table, names, inner=join_inner(cfg))


def join_inner(cfg):
    """(inner abi.Table, names) for the join configs, else None.  Built on the host; the library copies it per query."""
    if cfg not in ("c2join", "c2joins"):
        return None
    from heavydb_b200 import abi
    n = 10**5 if cfg == "c2join" else 10**4
    ids = np.arange(n, dtype=np.int32)
    x = ids.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    attr = ((x ^ (x >> np.uint64(31))) % np.uint64(1000)).astype(np.int32)
    t = abi.Table([(abi.kINT, True), (abi.kINT, True)])
    t.add_host_fragment([ids, attr])
    return t, ["id", "attr"]


def col_enc(col):
    return col[5] if len(col) > 5 else 0


def col_stride(col):
    return col[4] if len(col) > 4 else 1


def phys_type(col):
    """sql type of the PHYSICAL chunk elements (what the generator must write)."""
    from heavydb_b200 import abi
    e = col_enc(col)
    return {0: np_type(col[1]), 1: abi.kTINYINT, 2: abi.kSMALLINT, 4: abi.kINT}[e]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.load_start = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples inside [t0, t1] (the timed region); if the region is shorter than the sampling period, the
        samples of the whole loaded window (warm-up + timed steps) are used and the fact is reported."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [r for ts, r in self.rows if t0 is not None and t0 <= ts <= t1]
        window = "timed region"
        if len(inside) < 3:
            inside = [r for ts, r in self.rows if self.load_start is not None and self.load_start <= ts <= (t1 or ts)]
            window = "warm-up + timed region (timed region shorter than 3 sampling periods)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def np_type(t):
    from heavydb_b200 import abi
    return {"i64": abi.kBIGINT, "i32": abi.kINT, "f64": abi.kDOUBLE}[t]


def build_device_table(cfg, rows, frag_ids, torch):
    """Generate this rank's fragments directly in HBM with the counter-based generator (global row = frag_id*FRAG_ROWS+i)."""
    from heavydb_b200 import abi, executor
    cols, _, _, _ = CONFIGS[cfg]
    table = abi.Table([(np_type(c[1]), True) for c in cols], encoded_sizes=[col_enc(c) for c in cols])
    keep = []
    remaining = rows
    for fid in frag_ids:
        m = min(FRAG_ROWS, remaining)
        if m <= 0:
            break
        remaining -= m
        ptrs, stats = [], []
        for tag, col in enumerate(cols):
            _, t, lo, span = col[:4]
            stride = col_stride(col)
            ty = phys_type(col)
            buf = torch.empty(m * abi.SIZE_OF[ty], dtype=torch.uint8, device="cuda")
            executor.gen_column_device(buf.data_ptr(), ty, SEED, tag, fid * FRAG_ROWS, m, lo, span, stride=stride)
            keep.append(buf)
            ptrs.append(buf.data_ptr())
            st = abi.ChunkStats()
            if ty == abi.kDOUBLE:
                st.fp_min, st.fp_max = 0.0, 1.0
            else:
                st.int_min, st.int_max = lo, lo + (span - 1) * stride
            stats.append(st)
        table.add_device_fragment(m, ptrs, stats, fragment_id=fid)
    torch.cuda.synchronize()
    return table, keep


METRIC = "rows/sec and HBM GB/s on 1e9-row filter+groupby"   # BASELINE.json's metric; both arms print the same string


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the box's host cores, bounded sample."""
    import oracle_lib
    from heavydb_b200 import sqlmini
    from heavydb_b200 import abi
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    cols, sql, bpr, workload = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    frag_rows = 1 << 22  # 4 Mi rows per fragment, one fragment per thread (reference: one thread per fragment)
    nfrag = threads
    table = abi.Table([(np_type(c[1]), True) for c in cols], encoded_sizes=[col_enc(c) for c in cols])
    for f in range(nfrag):
        table.add_host_fragment([oracle_lib.gen_column(phys_type(c), SEED, tag, f * frag_rows, frag_rows, c[2], c[3], threads,
                                                       stride=col_stride(c)) for tag, c in enumerate(cols)])
    names = [c[0] for c in cols]
    unit = make_unit(args.config, sql, table, names)
    rows = nfrag * frag_rows
    guess = ENTRY_GUESS.get(args.config, 0)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = oracle_lib.execute(unit, table, entry_guess=guess, has_card=guess > 0, num_threads=threads)
        n_out = res.row_count()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = rows / (ms / 1e3)
    sample = f"{nfrag} fragments x {frag_rows} rows = {rows} rows of the same workload, one thread per fragment + host reduce"
    out = {
        "impl": "reference", "metric": METRIC,
        "note": "the reference's CPU algorithm for this path (restated in oracle/, validated against the reference's own tests) on the host cores",
        "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic", "config": {"workload": workload, "query": sql, "rows_per_step": rows, "groups_out": n_out},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)
    return 0


def cpu_baseline_sample(cfg, budget_s=15.0):
    import oracle_lib
    from heavydb_b200 import sqlmini
    from heavydb_b200 import abi
    cols, sql, _, _ = CONFIGS[cfg]
    threads = os.cpu_count() or 1
    frag_rows = 1 << 22
    nfrag = threads
    table = abi.Table([(np_type(c[1]), True) for c in cols], encoded_sizes=[col_enc(c) for c in cols])
    for f in range(nfrag):
        table.add_host_fragment([oracle_lib.gen_column(phys_type(c), SEED, tag, f * frag_rows, frag_rows, c[2], c[3], threads,
                                                       stride=col_stride(c)) for tag, c in enumerate(cols)])
    unit = make_unit(cfg, sql, table, [c[0] for c in cols])
    rows = nfrag * frag_rows
    guess = ENTRY_GUESS.get(cfg, 0)
    oracle_lib.execute(unit, table, entry_guess=guess, has_card=guess > 0, num_threads=threads)  # warm
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or (time.perf_counter() - t0 < budget_s and reps < 50):
        oracle_lib.execute(unit, table, entry_guess=guess, has_card=guess > 0, num_threads=threads)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": rows / dt, "unit": "rows/s", "cores": threads, "kind": "port",
            "sample": f"{nfrag} fragments x {frag_rows} rows ({rows} rows) of the same workload, one thread per fragment + host reduce, {reps} reps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows of the host-buffer end-to-end leg (0 = auto)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--force-kernel", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from heavydb_b200 import abi, build, executor, multigpu
    from heavydb_b200 import sqlmini
    build.build()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cols, sql, bytes_per_row, workload = CONFIGS[args.config]
    names = [c[0] for c in cols]
    rows = args.rows
    nfrag_per_rank = (rows + FRAG_ROWS - 1) // FRAG_ROWS
    frag_ids = multigpu.shard_fragments(range(nfrag_per_rank * world), rank, world)  # fragment_id % num_devices == rank
    table, keep = build_device_table(args.config, rows, frag_ids, torch)
    unit = make_unit(args.config, sql, table, names)
    ex = executor.Executor()
    eo = executor.execution_options(force_kernel=args.force_kernel)
    guess = ENTRY_GUESS.get(args.config, 0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    scan_ms, step_ms = [], []
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)  # let nvidia-smi come up so that samples exist for a millisecond-scale timed region
    result_rows = None
    t_begin = t_end = None
    sampler.load_start = time.time()
    for i in range(args.warmup + args.steps):
        barrier()
        if i == args.warmup:
            t_begin = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        part = ex.executePartial(guess, True, table, unit, eo=eo, has_cardinality_estimation=guess > 0, memory_level=abi.GPU_LEVEL)
        if dist is not None:
            multigpu.allreduce_partial(part, torch, dist)
        rs = part.finalize()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= args.warmup:
            step_ms.append(float(t.item()))
            scan_ms.append(part.kernel_ms())
        t_end = time.time()
        launches_per_step = rs.stats()["kernel_launches"]
        sort_us = rs.stats()["sort_us"]
        result_rows = rs.rowCount() if not sql.startswith("ESTIMATOR") else rs.getNDVEstimator()
        plan_kernel, plan_entries = int(rs.getQueryMemDesc().kernel), int(rs.getQueryMemDesc().entry_count)
        del rs, part
    clocks = sampler.stop(t_begin, t_end)
    ms = float(np.mean(step_ms))
    total_rows = rows * world
    value = total_rows / (ms / 1e3)
    k_ms = float(np.mean(scan_ms))
    peak, peak_src = measured_peak()
    achieved = rows * bytes_per_row / (k_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.config)
        if tr:
            traffic = (tr["dram_read_bytes"] + tr["dram_write_bytes"]) * (rows / tr["rows"])
            traffic_src = f"{tr['source']} (ncu --set full, {tr['rows']} rows/launch, scaled to {rows})"
    except Exception:
        pass
    out = {
        "metric": METRIC,
        "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic (counter-based splitmix64 columns generated in HBM; inputs 20 GB/GPU >> 126 MB L2, no flush needed)",
        "config": {"workload": workload, "query": sql, "rows_per_gpu": rows, "fragments_per_gpu": len(table.fragments),
                   "fragment_rows": FRAG_ROWS, "kernel": plan_kernel, "entry_count": plan_entries,
                   "groups_out": int(result_rows), "l2": "inputs larger than L2",
                   **({"sort_ms": sort_us / 1e3} if sort_us else {})},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": rows * bytes_per_row,
                     "peak_source": peak_src, "kernel": "b2q_k_scan", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_row": bytes_per_row},
        "clocks": clocks,
        "gpu_launches": int(launches_per_step) * args.steps,  # per step: b2q_k_init, b2q_k_scan (per launch), b2q_k_materialize
    }
    if rank == 0 and world == 1 and not args.no_e2e:
        out["e2e"] = e2e_leg(args, torch, ex, eo, cols, sql, names)
    elif world > 1:
        out["e2e"] = {"value": None, "unit": "rows/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                      "note": "host-buffer leg is measured at N=1"}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
            out["cpu_baseline"] = cpu_baseline_sample(args.config)
        except Exception as e:  # the oracle is only a reported baseline
            out["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def gpu_numa_cpus(torch):
    """CPUs of the NUMA node the GPU hangs off (so pinned host buffers are allocated next to its PCIe root)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None, node
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus, node
    except Exception as e:  # noqa
        return None, str(e)


def e2e_leg(args, torch, ex, eo, cols, sql, names):
    """Host (pinned) buffers -> C ABI -> host result.  H2D of every referenced column is inside the timed region."""
    import psutil
    cpus, node = gpu_numa_cpus(torch)
    old_aff = None
    if cpus:
        try:
            old_aff = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus & old_aff or cpus)
        except Exception:
            old_aff = None
    from heavydb_b200 import abi, executor
    from heavydb_b200 import sqlmini
    bytes_per_row = sum(abi.SIZE_OF[phys_type(c)] for c in cols)
    guess = ENTRY_GUESS.get(args.config, 0)
    avail = psutil.virtual_memory().available
    rows = args.e2e_rows or args.rows
    cap = int(avail * 0.4 // bytes_per_row)
    rows = max(FRAG_ROWS, min(rows, cap))
    table = abi.Table([(np_type(c[1]), True) for c in cols], encoded_sizes=[col_enc(c) for c in cols])
    keep = []
    for fi, b in enumerate(range(0, rows, FRAG_ROWS)):
        m = min(FRAG_ROWS, rows - b)
        harrs = []
        for tag, col in enumerate(cols):
            _, t, lo, span = col[:4]
            ty = phys_type(col)
            dev = torch.empty(m * abi.SIZE_OF[ty], dtype=torch.uint8, device="cuda")
            executor.gen_column_device(dev.data_ptr(), ty, SEED, tag, b, m, lo, span, stride=col_stride(col))
            host = torch.empty(m * abi.SIZE_OF[ty], dtype=torch.uint8, pin_memory=True)
            host.copy_(dev)
            keep.append(host)
            harrs.append(host.numpy().view(abi.NUMPY_OF[ty]))
            del dev
        fr = abi.Fragment(m, host_cols=harrs, stats=[], fragment_id=fi)
        for col in cols:
            _, t, lo, span = col[:4]
            st = abi.ChunkStats()
            if t == "f64":
                st.fp_min, st.fp_max = 0.0, 1.0
            else:
                st.int_min, st.int_max = lo, lo + (span - 1) * col_stride(col)
            fr.stats.append(st)
        table.fragments.append(fr)
    torch.cuda.synchronize()
    if old_aff:
        os.sched_setaffinity(0, old_aff)
    unit = make_unit(args.config, sql, table, names)
    bt = table.build(abi.CPU_LEVEL)
    # what this box's PCIe link delivers for plain copies out of the very same pinned buffers (no kernels, one stream):
    # the ceiling the end-to-end leg can reach — boxes of the pool differ by almost 2x here
    raw_gbs = None
    try:
        probe = [h for h in keep if h.numel() >= (64 << 20)]   # every buffer the timed leg will copy
        if probe:
            scratch = torch.empty(max(h.numel() for h in probe), dtype=torch.uint8, device="cuda")
            scratch[:probe[0].numel()].copy_(probe[0], non_blocking=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for h in probe:
                scratch[:h.numel()].copy_(h, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            raw_gbs = sum(h.numel() for h in probe) / (e0.elapsed_time(e1) / 1e3) / 1e9
            del scratch
    except Exception:  # the probe is informational
        raw_gbs = None
    times = []
    d2h = 0
    for i in range(2 + 3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rs = ex.executeWorkUnit(guess, True, bt, unit, eo=eo, has_cardinality_estimation=guess > 0, memory_level=abi.CPU_LEVEL)
        n = rs.rowCount()
        dt = time.perf_counter() - t0
        d2h = int(rs.getQueryMemDesc().buffer_size)
        phases = rs.stats()
        if i >= 2:
            times.append(dt)
        del rs
    dt = float(np.mean(times))
    return {"value": rows / dt, "unit": "rows/s", "h2d_bytes_per_step": int(rows * bytes_per_row), "d2h_bytes_per_step": d2h,
            "rows": rows, "ms_per_step": dt * 1e3, "host_memory": f"pinned, allocated on the GPU's NUMA node ({node})", "h2d_gbs": rows * bytes_per_row / dt / 1e9,
            "h2d_raw_gbs_same_buffers": raw_gbs, "groups_out": int(n),
            "host_phases_ms": {k[5:-3]: phases[k] / 1e3 for k in ("host_setup_us", "host_stream_us", "host_teardown_us")}}


if __name__ == "__main__":
    sys.exit(main())
