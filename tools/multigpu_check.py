#!/usr/bin/env python
"""Multi-GPU correctness check (one process per GPU, NCCL), run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

Every rank takes the fragments `fragment_id % world == rank` (InsertOrderFragmenter.cpp:435-443) and calls
b2q_execute_work_unit_dist: scan -> merge of the per-device tables by NCCL collectives INSIDE libb2q (all-reduce of the dense
arrays; all-gather + re-probe for baseline hash; all-gather + OR for estimator bitmaps — replacing the host-side
reduceMultiDeviceResults) -> materialise incl. ORDER BY / LIMIT and the join level; rank 0 compares the result with the
oracle run over the WHOLE table.  `--merge python` keeps the older harness (b2q_execute_partial + torch.distributed
all-reduce per array + finalize) for the dense layouts.  Test infrastructure: the oracle is the checker only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpu_util as gu  # noqa: E402
import join_tables as jt  # noqa: E402
import oracle_lib  # noqa: E402
import order_queries as oq  # noqa: E402
from heavydb_b200 import abi, executor, multigpu, sqlmini  # noqa: E402
from test_gpu_parity import RAND_NAMES, RAND_QUERIES, random_table  # noqa: E402


def shard(table, rank, world):
    """The rank's fragments as a device-resident table (fragment ids kept)."""
    sub = abi.Table(table.col_types, encoded_sizes=table.encoded_sizes, deleted_column=table.deleted_column)
    keep = []
    for f in table.fragments:
        if f.fragment_id % world != rank:
            continue
        ptrs = []
        for a in f.host_cols:
            t = torch.from_numpy(a.view(np.uint8).copy()).cuda()
            keep.append(t)
            ptrs.append(t.data_ptr())
        sub.add_device_fragment(f.num_tuples, ptrs, f.stats, fragment_id=f.fragment_id)
    return sub, keep


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ex = executor.Executor()
    lib_merge = "--merge" not in sys.argv or sys.argv[sys.argv.index("--merge") + 1] != "python"
    comm = None
    if lib_merge:   # the 128-byte NCCL id travels over the process group the launcher already gave us
        box = [executor.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = executor.Comm.init_rank(box[0], world, rank, device=local)
    checked = 0
    cases = []
    table = random_table(200000, seed=17, frag_rows=12500)    # 16 fragments
    for sql in RAND_QUERIES + oq.RAND_ORDER_QUERIES:
        cases.append((sql, table, RAND_NAMES, None))
    fact, dim = jt.fact_table(150000, seed=5, frag_rows=10000), jt.dim_table()
    for sql in jt.JOIN_QUERIES:
        cases.append((sql, fact, jt.FACT_NAMES, (dim, jt.DIM_NAMES)))
    shards = {}
    for sql, tbl, names, inner in cases:
        unit = sqlmini.parse(sql, tbl, names, inner=inner)
        # the plan is made from the WHOLE table's chunk stats on every rank (same ranges => position-aligned tables)
        if id(tbl) not in shards:
            shards[id(tbl)] = shard(tbl, rank, world)
        sub, _ = shards[id(tbl)]
        plan = ex.plan(unit, tbl, max_groups_buffer_entry_guess=4000, has_cardinality_estimation=True)
        if plan.query_desc_type == abi.GroupByBaselineHash and not lib_merge:
            continue    # not position-aligned across devices: only the library's re-probe merge handles it
        # the other ranks' fragments travel as chunk stats only, so that every rank derives the same key ranges
        view = abi.Table(tbl.col_types, encoded_sizes=tbl.encoded_sizes, deleted_column=tbl.deleted_column)
        for f in sub.fragments:
            view.fragments.append(f)
        for f in tbl.fragments:
            if f.fragment_id % world != rank:
                view.add_remote_fragment(f.num_tuples, f.stats, f.fragment_id)
        part = None
        if lib_merge:
            rs = executor.execute_work_unit_dist(comm, ex, 4000, True, view, unit, has_cardinality_estimation=True)
        else:
            part = ex.executePartial(4000, True, view, unit, has_cardinality_estimation=True, memory_level=abi.GPU_LEVEL)
            multigpu.allreduce_partial(part, torch, dist)
            rs = part.finalize()
        rows = rs.rows()
        if rank == 0:
            ref = oracle_lib.execute(unit, tbl, entry_guess=4000, has_card=True, num_threads=8)
            want = ref.rows()
            if unit.unit.num_order_entries:
                gu.rows_equal_ordered(rows, want)
            else:
                gu.rows_equal(rows, want, col_tol=gu.column_tolerances(ref.plan, sum(f.num_tuples for f in tbl.fragments)))
            assert rs.rowCount() == ref.row_count(), sql
        checked += 1
        del rs, part
    # the estimator query: per-rank bitmaps are OR-merged (all-gather + OR; reduce_estimator_results)
    for cols in (["k32"], ["k16", "nn32"], ["sparse"]):
        b = abi.UnitBuilder(table)
        b.estimator([RAND_NAMES.index(c) for c in cols])
        unit = b.build()
        sub, _ = shards[id(table)]
        view = abi.Table(table.col_types)
        for f in sub.fragments:
            view.fragments.append(f)
        for f in table.fragments:
            if f.fragment_id % world != rank:
                view.add_remote_fragment(f.num_tuples, f.stats, f.fragment_id)
        part = None
        if lib_merge:
            rs = executor.execute_work_unit_dist(comm, ex, 1, True, view, unit)
        else:
            part = ex.executePartial(1, True, view, unit, memory_level=abi.GPU_LEVEL)
            multigpu.allreduce_partial(part, torch, dist)
            rs = part.finalize()
        if rank == 0:
            ref = oracle_lib.execute(unit, table, num_threads=8)
            assert np.array_equal(rs.getHostEstimatorBuffer(), ref.buffer().view(np.uint8)), cols
            assert rs.getNDVEstimator() == ref.ndv_estimator()
        checked += 1
        del rs, part
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: {checked} queries, world={world}, merge={'libb2q/NCCL' if lib_merge else 'python'}", flush=True)
    if comm is not None:
        comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
