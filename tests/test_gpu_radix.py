"""Baseline-hash GROUP BY through the radix-partitioned aggregation (heavydb_b200/csrc/radix_agg.cu): CUDA path vs the
oracle on shapes that exercise every branch — several partitions, one and many tuple words, region overflow (skewed keys
inserted straight into the HBM table), probe clusters that run off a slice into the overflow areas, listed raw tuples,
host-resident tables (one launch per slice over the same buffers) — and the per-row probe kernel it replaces."""
import numpy as np
import pytest

import gpu_util as gu
import oracle_lib
import sqlmini
from heavydb_b200 import abi, executor

pytestmark = pytest.mark.gpu


def sparse_table(n, ndv, seed, frag_rows, skew=0.0, nullable_v=False, key32=False):
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, ndv, n)
    if skew:
        ids[rng.random(n) < skew] = 7
    if key32:
        key = (ids.astype(np.int64) * 21001 - 10**9).astype(np.int32)    # range too wide for a perfect hash
    else:
        key = ((ids.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(1)).astype(np.int64)
    v = rng.integers(-10**6, 10**6, n).astype(np.int64)
    if nullable_v:
        v[rng.random(n) < 0.2] = abi.NULL_OF[abi.kBIGINT]
    w = rng.integers(-100, 100, n).astype(np.int16)
    d = rng.normal(0, 100, n)
    d[rng.random(n) < 0.1] = abi.NULL_OF[abi.kDOUBLE]
    f = rng.integers(0, 1000, n).astype(np.int32)
    t = abi.Table([(abi.kINT if key32 else abi.kBIGINT, True), (abi.kBIGINT, not nullable_v), (abi.kSMALLINT, True), (abi.kDOUBLE, False), (abi.kINT, True)])
    for b in range(0, n, frag_rows):
        t.add_host_fragment([key[b:b + frag_rows], v[b:b + frag_rows], w[b:b + frag_rows], d[b:b + frag_rows], f[b:b + frag_rows]])
    return t, ["key", "v", "w", "d", "f"], len(np.unique(key))


QUERIES = [
    "SELECT key, SUM(v) FROM t GROUP BY key;",                                            # 16-byte tuples: one vector store / load
    "SELECT key, COUNT(*) FROM t WHERE f < 500 GROUP BY key;",                            # key-only tuples
    "SELECT key, SUM(v), COUNT(v), MIN(w), MAX(d), AVG(d), COUNT(*) FROM t WHERE f >= 100 GROUP BY key;",   # 4 tuple words, every accumulator kind
]


@pytest.mark.parametrize("n,ndv,frag_rows", [(50_000, 30_000, 20_000), (400_000, 100_000, 150_000), (1_500_000, 400_000, 1 << 19)])
@pytest.mark.parametrize("nullable_v", [False, True])
def test_radix_matches_oracle(n, ndv, frag_rows, nullable_v):
    table, names, real_ndv = sparse_table(n, ndv, seed=n + nullable_v, frag_rows=frag_rows, nullable_v=nullable_v)
    dev = gu.DeviceTable(table)
    for sql in QUERIES:
        unit = sqlmini.parse(sql, table, names)
        rs, _ = gu.run_both(unit, table, entry_guess=int(real_ndv * 1.5), has_card=True, dev_table=dev, oracle_threads=4)
        assert rs.getQueryMemDesc().kernel == abi.KERNEL_BASELINE_GLOBAL
        assert rs.stats()["kernel_launches"] >= 4          # init + 2 radix passes + materialise


def test_radix_key32_and_host_resident():
    table, names, ndv = sparse_table(300_000, 50_000, seed=3, frag_rows=70_000, key32=True)
    for sql in QUERIES:
        unit = sqlmini.parse(sql, table, names)
        gu.run_both(unit, table, entry_guess=int(ndv * 1.5), has_card=True, oracle_threads=4)
        gu.run_both(unit, table, entry_guess=int(ndv * 1.5), has_card=True, device_resident=False, oracle_threads=4)


def test_radix_skewed_keys_overflow_their_regions():
    """60 % of the rows carry ONE key: its partition's regions fill up and the rest goes straight into the HBM table."""
    table, names, ndv = sparse_table(1_000_000, 200_000, seed=11, frag_rows=1 << 18, skew=0.6)
    dev = gu.DeviceTable(table)
    for sql in QUERIES:
        gu.run_both(sqlmini.parse(sql, table, names), table, entry_guess=int(ndv * 1.5), has_card=True, dev_table=dev, oracle_threads=4)


def test_radix_nearly_full_table_uses_overflow_areas():
    """entry_count barely above the number of keys: long probe clusters cross slice ends (overflow areas, listed tuples)."""
    table, names, ndv = sparse_table(600_000, 120_000, seed=21, frag_rows=1 << 18)
    dev = gu.DeviceTable(table)
    for guess in (ndv + 1, int(ndv * 1.02), int(ndv * 1.1)):
        for sql in QUERIES[:1] + QUERIES[2:]:
            gu.run_both(sqlmini.parse(sql, table, names), table, entry_guess=guess, has_card=True, dev_table=dev, oracle_threads=4)


def test_probe_kernel_still_agrees():
    table, names, ndv = sparse_table(200_000, 60_000, seed=5, frag_rows=80_000, nullable_v=True)
    dev = gu.DeviceTable(table)
    for sql in QUERIES:
        unit = sqlmini.parse(sql, table, names)
        rs, _ = gu.run_both(unit, table, entry_guess=int(ndv * 1.5), has_card=True, dev_table=dev, force_kernel=abi.KERNEL_BASELINE_PROBE, oracle_threads=4)
        assert rs.stats()["kernel_launches"] == 3           # init + scan + materialise


def test_radix_out_of_slots_and_empty_input():
    table, names, ndv = sparse_table(20_000, 5_000, seed=8, frag_rows=20_000)
    unit = sqlmini.parse(QUERIES[0], table, names)
    with pytest.raises(executor.QueryExecutionError) as ei:
        executor.Executor().executeWorkUnit(ndv // 2, True, table, unit, has_cardinality_estimation=True)
    assert ei.value.code == abi.ERR_OUT_OF_SLOTS
    empty, names, _ = sparse_table(0, 1, seed=1, frag_rows=10)
    empty.add_host_fragment([np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int16), np.zeros(0), np.zeros(0, dtype=np.int32)])
    rs = executor.Executor().executeWorkUnit(64, True, empty, sqlmini.parse(QUERIES[0], empty, names), has_cardinality_estimation=True)
    assert rs.rowCount() == 0
