"""`-m gpu` parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (task statement / north_star): bit-exact COUNT / SUM(int64) / MIN / MAX and integer layouts; SUM/AVG(double)
within 1e-6 relative.  Covers the reference's golden table (Tests/ExecuteTest.cpp), every kernel family, every
column type, NULL keys / NULL arguments, empty and ragged fragments, host-resident (H2D inside the call) and
HBM-resident inputs, and BASELINE.json's configs at reduced row counts.
"""
import numpy as np
import pytest

import gpu_util as gu
import oracle_lib
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_oracle_golden import COUNT_DISTINCT_QUERIES, FLOAT_QUERIES, MULTI_KEY_QUERIES, NULL_LOGIC_QUERIES, PATH_QUERIES, REFERENCE_QUERIES
from test_planner_parity import EXTRA

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    table = rt.make_table(rt.test_rows())
    return table, gu.DeviceTable(table)


@pytest.mark.parametrize("sql", REFERENCE_QUERIES + PATH_QUERIES + EXTRA + MULTI_KEY_QUERIES + NULL_LOGIC_QUERIES + COUNT_DISTINCT_QUERIES + FLOAT_QUERIES)
def test_golden_table_device_resident(golden, sql):
    table, dev = golden
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    gu.run_both(unit, table, entry_guess=48, has_card=True, dev_table=dev)


@pytest.mark.parametrize("sql", (REFERENCE_QUERIES + PATH_QUERIES)[::3])
def test_golden_table_host_resident(golden, sql):
    """Same queries with HOST column buffers: the call stages them to the GPU itself (fetchChunks' H2D)."""
    table, _ = golden
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    gu.run_both(unit, table, entry_guess=48, has_card=True, device_resident=False)


@pytest.mark.parametrize("sql", [q for q in PATH_QUERIES if "GROUP BY" in q][::2])
def test_golden_table_forced_global_kernel(golden, sql):
    """Perfect-hash queries forced onto the HBM/L2 table kernel (the path tables too large for smem take)."""
    table, dev = golden
    unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
    p = executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=48, has_cardinality_estimation=True)
    if p.query_desc_type != abi.GroupByPerfectHash:
        pytest.skip("not a perfect-hash plan")
    gu.run_both(unit, table, entry_guess=48, has_card=True, force_kernel=abi.KERNEL_PERFECT_GLOBAL, dev_table=dev)


@pytest.mark.parametrize("bigint_count", [False, True])
def test_bigint_count(golden, bigint_count):
    table, dev = golden
    for sql in ["SELECT x, COUNT(*), COUNT(ofq), COUNT(dn) FROM test GROUP BY x;", "SELECT COUNT(*), COUNT(t) FROM test WHERE x = 7;",
                "SELECT y, COUNT(*) FROM test GROUP BY y;"]:
        unit = sqlmini.parse(sql, table, rt.TEST_NAMES, bigint_count=bigint_count)
        gu.run_both(unit, table, bigint_count=bigint_count, dev_table=dev)


def test_empty_table_and_empty_fragments():
    empty = rt.make_table([])
    for sql in ["SELECT COUNT(*) FROM test;", "SELECT SUM(x), MIN(y), MAX(t), AVG(d), COUNT(z) FROM test;"]:
        unit = sqlmini.parse(sql, empty, rt.TEST_NAMES)
        gu.run_both(unit, empty, device_resident=False)
    unit = sqlmini.parse("SELECT x, COUNT(*) FROM test GROUP BY x;", empty, rt.TEST_NAMES)
    with pytest.raises(executor.CardinalityEstimationRequired):
        executor.Executor().executeWorkUnit(0, True, empty, unit)
    rs, _ = gu.run_both(unit, empty, entry_guess=16, has_card=True, device_resident=False)
    assert rs.rowCount() == 0 and rs.isEmpty()
    # a table whose first / middle / last fragments are empty
    rows = rt.test_rows()
    cols = rt.to_columns(rows)
    t = abi.Table([(ty, nn) for _, ty, nn in rt.TEST_COLS])
    cuts = [0, 0, 3, 3, 11, 20, 20]
    for a, b in zip(cuts[:-1], cuts[1:]):
        t.add_host_fragment([c[a:b] for c in cols])
    for sql in ["SELECT x, COUNT(*), SUM(t) FROM test GROUP BY x;", "SELECT SUM(t), COUNT(*) FROM test WHERE z > 0;"]:
        gu.run_both(sqlmini.parse(sql, t, rt.TEST_NAMES), t)


# ---- randomized tables -------------------------------------------------------------------------------------
def _columnar_or_both_refuse(unit, table, dev, **kw):
    """Run with the columnar hint; layouts the reference's reader cannot read (see planner.cpp) are refused by both."""
    try:
        oracle_lib.plan(unit, table, entry_guess=kw.get("entry_guess", 0), has_card=kw.get("has_card", False), output_columnar=True)
    except oracle_lib.OracleError as e:
        assert e.code == abi.ERR_UNSUPPORTED
        with pytest.raises(executor.UnsupportedOnThisPath):
            executor.Executor().plan(unit, table, eo=executor.execution_options(output_columnar_hint=True),
                                     max_groups_buffer_entry_guess=kw.get("entry_guess", 0),
                                     has_cardinality_estimation=kw.get("has_card", False))
        return False
    rs, _ = gu.run_both(unit, table, dev_table=dev, output_columnar=True, **kw)
    assert rs.getQueryMemDesc().output_columnar == 1
    return True


def test_golden_table_columnar_output():
    """eo.output_columnar_hint (--enable-columnar-output / the columnar hint): ResultSet.h:72-84 layout, bit-exact."""
    table = rt.make_table(rt.test_rows())
    dev = gu.DeviceTable(table)
    ran = 0
    for sql in REFERENCE_QUERIES + PATH_QUERIES + EXTRA + MULTI_KEY_QUERIES:
        unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
        ran += _columnar_or_both_refuse(unit, table, dev, entry_guess=48, has_card=True)
    assert ran >= 40


def test_random_table_columnar_output():
    table = random_table(50000, seed=77, frag_rows=16384)
    dev = gu.DeviceTable(table)
    ran = 0
    for sql in RAND_QUERIES:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        ran += _columnar_or_both_refuse(unit, table, dev, entry_guess=3001, has_card=True)
    assert ran >= 15


RAND_COLS = [
    ("k8", abi.kTINYINT, False), ("k16", abi.kSMALLINT, False), ("k32", abi.kINT, False), ("k64", abi.kBIGINT, False),
    ("nn32", abi.kINT, True), ("nn64", abi.kBIGINT, True), ("a8", abi.kTINYINT, False), ("a16", abi.kSMALLINT, True),
    ("a32", abi.kINT, False), ("a64", abi.kBIGINT, False), ("big", abi.kBIGINT, True), ("d", abi.kDOUBLE, False),
    ("dnn", abi.kDOUBLE, True), ("sparse", abi.kBIGINT, True), ("f32", abi.kFLOAT, False), ("fnn", abi.kFLOAT, True),
]
RAND_NAMES = [c[0] for c in RAND_COLS]


def random_table(n, seed, frag_rows):
    rng = np.random.default_rng(seed)

    def with_nulls(a, t, p=0.1):
        if n:
            a = a.copy()
            a[rng.random(n) < p] = abi.NULL_OF[t]
        return a
    cols = [
        with_nulls(rng.integers(-5, 20, n).astype(np.int8), abi.kTINYINT),
        with_nulls(rng.integers(100, 400, n).astype(np.int16), abi.kSMALLINT),
        with_nulls(rng.integers(-1000, 1000, n).astype(np.int32), abi.kINT),
        with_nulls(rng.integers(10**9, 10**9 + 5000, n).astype(np.int64), abi.kBIGINT),
        rng.integers(0, 300, n).astype(np.int32),
        rng.integers(-50, 50, n).astype(np.int64),
        with_nulls(rng.integers(-128 + 1, 128, n).astype(np.int8), abi.kTINYINT),
        rng.integers(-30000, 30000, n).astype(np.int16),
        with_nulls(rng.integers(-2**31 + 1, 2**31, n).astype(np.int32), abi.kINT),
        with_nulls(rng.integers(-2**40, 2**40, n).astype(np.int64), abi.kBIGINT),
        rng.integers(-2**62, 2**62, n).astype(np.int64),           # sums wrap around: exercises the lo/hi carry path
        with_nulls(rng.normal(0, 1e3, n), abi.kDOUBLE),
        rng.random(n),
        rng.integers(0, 2000, n).astype(np.int64) * 7919 * 10**9,  # sparse keys -> baseline hash
        with_nulls(rng.normal(0, 50, n).astype(np.float32), abi.kFLOAT),        # FLOAT chunks: 4 bytes, NULL_FLOAT = FLT_MIN
        (rng.random(n) * 8 - 1).astype(np.float32),
    ]
    t = abi.Table([(ty, nn) for _, ty, nn in RAND_COLS])
    if n == 0:
        t.add_host_fragment(cols)
    for b in range(0, n, frag_rows):
        t.add_host_fragment([c[b:b + frag_rows] for c in cols])
    return t


RAND_QUERIES = [
    "SELECT COUNT(*) FROM r WHERE a32 < 0;",
    "SELECT COUNT(*), COUNT(a8), COUNT(a64), COUNT(d), SUM(a8), SUM(a16), SUM(a32), SUM(a64), SUM(big) FROM r;",
    "SELECT MIN(a8), MAX(a8), MIN(a16), MAX(a16), MIN(a32), MAX(a32), MIN(a64), MAX(a64), MIN(big), MAX(big) FROM r WHERE nn32 >= 10;",
    "SELECT SUM(d), AVG(d), MIN(d), MAX(d), SUM(dnn), AVG(dnn), MIN(dnn), MAX(dnn) FROM r WHERE dnn < 0.75;",
    "SELECT AVG(a8), AVG(a16), AVG(a32), AVG(a64) FROM r WHERE d > -100.5 OR a16 < 0;",
    "SELECT nn32, COUNT(*), SUM(a64), SUM(big) FROM r GROUP BY nn32;",
    "SELECT nn32, SUM(a32), MIN(a32), MAX(a32), AVG(a32), COUNT(a32) FROM r WHERE nn64 < 25 GROUP BY nn32;",
    "SELECT nn64, SUM(d), AVG(d), MIN(d), MAX(d), COUNT(d) FROM r GROUP BY nn64;",
    "SELECT nn64, SUM(dnn), AVG(dnn), MIN(dnn), MAX(dnn) FROM r WHERE a8 <> 3 GROUP BY nn64;",
    "SELECT k8, COUNT(*), SUM(a16) FROM r GROUP BY k8;",
    "SELECT k16, COUNT(*), MIN(a64), MAX(a64), SUM(a64) FROM r WHERE k16 > 150 GROUP BY k16;",
    "SELECT k32, COUNT(*), AVG(a16), SUM(big) FROM r WHERE nn32 < 200 AND (a16 > 0 OR nn64 = 7) GROUP BY k32;",
    "SELECT k64, COUNT(*), SUM(a32) FROM r GROUP BY k64;",
    "SELECT k64, SUM(a64) FROM r WHERE k64 >= 1000002000 GROUP BY k64;",
    "SELECT nn32, MIN(a8), MAX(a16), MIN(big), MAX(big) FROM r GROUP BY nn32;",
    "SELECT nn32, SUM(a16) FROM r GROUP BY nn32;",
    "SELECT sparse, COUNT(*), SUM(a64), MIN(a32), MAX(d), AVG(dnn) FROM r GROUP BY sparse;",
    "SELECT sparse, SUM(big) FROM r WHERE nn32 < 150 GROUP BY sparse;",
    "SELECT COUNT(*) FROM r WHERE (nn32 < 100 AND nn64 > 0) OR (a16 >= 100 AND a16 <= 20000 AND dnn <> 0.5);",
    "SELECT COUNT(*), SUM(nn64) FROM r WHERE a64 > 5.5;",
    "SELECT COUNT(*), SUM(nn64) FROM r WHERE d <= 10;",
    "SELECT k8, nn64, COUNT(*), SUM(a32), MIN(big), AVG(d) FROM r GROUP BY k8, nn64;",            # multi-column keys
    "SELECT nn32, k8, nn64, COUNT(*), SUM(a64) FROM r WHERE nn32 < 30 GROUP BY nn32, k8, nn64;",
    "SELECT k32, k8, SUM(dnn), COUNT(a16) FROM r WHERE k32 >= -20 AND k32 < 20 GROUP BY k32, k8;",
    "SELECT k8, nn64, MIN(d), MAX(d) FROM r GROUP BY k8, nn64;",                                   # keyed (not keyless) layouts
    "SELECT k8, k16, SUM(a64) FROM r GROUP BY k8, k16;",
    "SELECT nn64, k8, AVG(d) FROM r WHERE a32 > 0 GROUP BY nn64, k8;",
    "SELECT k8, COUNT(*), SUM(a64) FROM r WHERE a64 IS NOT NULL AND NOT (k16 IS NULL OR a8 IN (1, 2, 3, -4)) GROUP BY k8;",   # NOT / IS NULL / IN
    "SELECT COUNT(*), COUNT(d), MIN(a32) FROM r WHERE d IS NULL OR NOT (a16 BETWEEN -1000 AND 1000 AND dnn < 0.5) OR big IS NULL;",
    "SELECT nn32, COUNT(*) FROM r WHERE NOT (NOT (k32 >= 0)) AND k64 NOT IN (1000000001, 1000000002) AND nn64 IS NOT NULL GROUP BY nn32;",
    "SELECT k8, COUNT(*), SUM(a16) FROM r WHERE a8 < k8 OR a32 > a64 GROUP BY k8;",                       # column OP column: int8/int8, int32/int64
    "SELECT COUNT(*), MIN(d), MAX(dnn) FROM r WHERE d <= dnn AND NOT (a16 = nn32) AND (k16 >= nn32 OR big < a64);",   # double/double, int16/int32
    "SELECT nn64, COUNT(*) FROM r WHERE d > a32 AND nn32 <> nn64 GROUP BY nn64;",                       # double vs int32
    # FLOAT arguments / filters: widened to double on load, 4-byte float images in the slots
    "SELECT k8, COUNT(f32), MIN(f32), MAX(f32), SUM(fnn), AVG(f32) FROM r WHERE fnn > 0.5 GROUP BY k8;",
    "SELECT COUNT(*), COUNT(f32), MIN(fnn), MAX(fnn), SUM(f32), AVG(fnn) FROM r WHERE f32 < 10 OR f32 IS NULL;",
    "SELECT nn32, MAX(f32), MIN(f32) FROM r WHERE f32 <= d AND fnn <> 3 GROUP BY nn32;",
    "SELECT sparse, SUM(fnn), MAX(f32) FROM r GROUP BY sparse;",
    # COUNT(DISTINCT): per-group bitmaps in HBM; shared-memory, HBM/L2, non-grouped and baseline-hash group tables
    "SELECT k8, COUNT(DISTINCT a16), COUNT(DISTINCT nn32), COUNT(*) FROM r GROUP BY k8;",
    "SELECT COUNT(DISTINCT a32), COUNT(DISTINCT k64), COUNT(DISTINCT a8), SUM(a16) FROM r WHERE nn32 < 250;",
    "SELECT nn32, COUNT(DISTINCT k16), AVG(d) FROM r WHERE a16 BETWEEN -20000 AND 20000 AND k16 > 120 GROUP BY nn32;",
    "SELECT sparse, COUNT(DISTINCT nn64), SUM(a32) FROM r GROUP BY sparse;",
    "SELECT k8, nn64, COUNT(DISTINCT a8), MIN(big) FROM r GROUP BY k8, nn64;",
]


@pytest.mark.parametrize("n,frag_rows", [(1, 10), (31, 7), (1000, 333), (4097, 4097), (50000, 16384), (300000, 100000)])
def test_random_tables(n, frag_rows):
    table = random_table(n, seed=n, frag_rows=frag_rows)
    dev = gu.DeviceTable(table)
    for sql in RAND_QUERIES:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            oracle_lib.plan(unit, table, entry_guess=3001, has_card=True)
        except oracle_lib.OracleError as e:   # e.g. an all-NULL key column of a composite key: both sides must refuse
            assert e.code == abi.ERR_UNSUPPORTED
            with pytest.raises(executor.UnsupportedOnThisPath):
                executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=3001, has_cardinality_estimation=True)
            continue
        gu.run_both(unit, table, entry_guess=3001, has_card=True, dev_table=dev)


def test_random_table_forced_global_and_host():
    table = random_table(20000, seed=5, frag_rows=6000)
    dev = gu.DeviceTable(table)
    for sql in RAND_QUERIES:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        p = executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=3001, has_cardinality_estimation=True)
        if p.query_desc_type == abi.GroupByPerfectHash:
            gu.run_both(unit, table, force_kernel=abi.KERNEL_PERFECT_GLOBAL, dev_table=dev)
        gu.run_both(unit, table, entry_guess=3001, has_card=True, device_resident=False)


def test_plain_word_layout_of_the_hbm_table_kernels():
    """B2Q_GLOBAL_SPLIT=0 (read once per process, hence the child): COUNT / integer SUM of the HBM-table kernels as plain int64
    words updated with RED.ADD.64, the touched flag derived at materialise (flag | accumulator != 0) — same rows, same buffers."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B2Q_GLOBAL_SPLIT="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu",
                        "tests/test_gpu_parity.py::test_random_table_forced_global_and_host",
                        "tests/test_gpu_parity.py::test_touched_flag_cases_of_the_global_kernels"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]


def test_touched_flag_cases_of_the_global_kernels():
    """Keyed perfect-hash groups whose SUM is zero (all-zero values, cancelling values, a single zero) must still come out as
    groups, and untouched entries must not: the HBM-table kernels see them through the touched flag."""
    n = 4000
    rng = np.random.default_rng(77)
    k = rng.integers(0, 50, n).astype(np.int64)
    v = rng.integers(-3, 4, n).astype(np.int64)
    v[k == 7] = 0                      # group 7: only zeros
    v[k == 9] = np.where(np.arange((k == 9).sum()) % 2 == 0, 5, -5)[: (k == 9).sum()]     # group 9: cancels (or +5)
    keep = (k != 11) & (k != 12)       # groups 11 / 12 never appear: empty entries inside the range
    k, v = k[keep], v[keep]
    t = abi.Table([(abi.kBIGINT, True), (abi.kBIGINT, True)])
    t.add_host_fragment([k, v])
    for sql in ("SELECT k, SUM(v) FROM t GROUP BY k;", "SELECT k, SUM(v), MIN(v) FROM t GROUP BY k;",
                "SELECT k, SUM(v) FROM t WHERE v <= 0 GROUP BY k;"):
        unit = sqlmini.parse(sql, t, ["k", "v"])
        p = executor.Executor().plan(unit, t)
        assert p.query_desc_type == abi.GroupByPerfectHash and (not p.keyless_hash or "MIN" in sql)   # MIN(v) makes a keyless layout
        gu.run_both(unit, t, force_kernel=abi.KERNEL_PERFECT_GLOBAL)
        gu.run_both(unit, t)


def test_baseline_out_of_slots():
    """Fewer entries than distinct keys: the reference returns -pos => OUT_OF_SLOTS (GroupByAndAggregate.cpp:1149-1154)."""
    table = random_table(5000, seed=9, frag_rows=5000)
    unit = sqlmini.parse("SELECT sparse, COUNT(*) FROM r GROUP BY sparse;", table, RAND_NAMES)
    with pytest.raises(executor.QueryExecutionError) as ei:
        executor.Executor().executeWorkUnit(100, True, table, unit, has_cardinality_estimation=True)
    assert ei.value.code == abi.ERR_OUT_OF_SLOTS
    with pytest.raises(oracle_lib.OracleError) as oe:
        oracle_lib.execute(unit, table, entry_guess=100, has_card=True)
    assert oe.value.code < 0  # -pos


def test_stale_chunk_stats_are_detected():
    """A key outside the advertised chunk range would make the reference write out of bounds; we flag it."""
    t = abi.Table([(abi.kINT, True), (abi.kBIGINT, True)])
    t.add_host_fragment([np.array([1, 2, 3, 50], dtype=np.int32), np.arange(4, dtype=np.int64)])
    t.fragments[0].stats[0].int_max = 3   # lie about the max
    unit = sqlmini.parse("SELECT g, SUM(v) FROM t GROUP BY g;", t, ["g", "v"])
    with pytest.raises(executor.QueryExecutionError) as ei:
        executor.Executor().executeWorkUnit(0, True, t, unit)
    assert ei.value.code == abi.ERR_KEY_OUT_OF_RANGE


# ---- BASELINE.json configs at oracle-friendly sizes (seeded counter-based generator on both sides) ----------
SEED = 0x5EED


def gen_table(n, spec, frag_rows):
    """spec: [(name, sql_type, col_tag, lo, span)], all NOT NULL.  Host copy via the oracle's generator."""
    t = abi.Table([(ty, True) for _, ty, _, _, _ in spec])
    for fi, b in enumerate(range(0, n, frag_rows)):
        m = min(frag_rows, n - b)
        t.add_host_fragment([oracle_lib.gen_column(ty, SEED, tag, b, m, lo, span) for _, ty, tag, lo, span in spec])
    return t, [s[0] for s in spec]


def test_device_generator_matches_oracle_generator():
    import torch
    n = 100003
    for ty, lo, span in [(abi.kBIGINT, 0, 10**6), (abi.kINT, -5, 10**4), (abi.kSMALLINT, 0, 1000), (abi.kTINYINT, -3, 100),
                         (abi.kDOUBLE, 0, 1), (abi.kBIGINT, -2**40, 2**41)]:
        want = oracle_lib.gen_column(ty, SEED, 3, 12345, n, lo, span)
        buf = torch.empty(n * abi.SIZE_OF[ty], dtype=torch.uint8, device="cuda")
        executor.gen_column_device(buf.data_ptr(), ty, SEED, 3, 12345, n, lo, span)
        torch.cuda.synchronize()
        got = buf.cpu().numpy().view(abi.NUMPY_OF[ty])
        assert np.array_equal(got, want)


C2_SPEC = [("c0", abi.kBIGINT, 0, 0, 10**6), ("c1", abi.kBIGINT, 1, 0, 10**6), ("c2", abi.kBIGINT, 2, 0, 10**6),
           ("c3", abi.kBIGINT, 3, 0, 10**6), ("g", abi.kINT, 4, 0, 10**4)]


@pytest.mark.parametrize("k", [10000, 500000, 10**6])  # ~1 %, 50 %, 100 % selectivity
def test_config2_filter_groupby_1e4_groups(k):
    table, names = gen_table(3_000_000, C2_SPEC, frag_rows=1 << 20)
    dev = gu.DeviceTable(table)
    for sql in [f"SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < {k} GROUP BY g;",
                f"SELECT g, SUM(c1), SUM(c2), SUM(c3), COUNT(*) FROM t WHERE c0 < {k} GROUP BY g;"]:
        rs, _ = gu.run_both(sqlmini.parse(sql, table, names), table, dev_table=dev, oracle_threads=8)
        assert rs.getQueryMemDesc().kernel == abi.KERNEL_PERFECT_SMEM


def test_config1_count_filter_int32():
    table, names = gen_table(1_000_000, [("a", abi.kINT, 0, 0, 10**6)], frag_rows=1_000_000)
    rs, ref = gu.run_both(sqlmini.parse("SELECT COUNT(*) FROM t WHERE a < 500000;", table, names), table)
    assert rs.getQueryMemDesc().kernel == abi.KERNEL_NON_GROUPED


def test_config3_low_card_avg_double():
    spec = [("f", abi.kBIGINT, 0, 0, 10**6), ("g", abi.kINT, 1, 0, 256), ("v", abi.kDOUBLE, 2, 0, 1)]
    table, names = gen_table(3_000_000, spec, frag_rows=1 << 20)
    rs, _ = gu.run_both(sqlmini.parse("SELECT g, AVG(v) FROM t WHERE f < 500000 GROUP BY g;", table, names), table, oracle_threads=8)
    assert rs.getQueryMemDesc().kernel == abi.KERNEL_PERFECT_SMEM and rs.rowCount() == 256


def test_config4_high_card_dense_and_sparse():
    n, card = 4_000_000, 10**6
    spec = [("key", abi.kBIGINT, 0, 0, card), ("v", abi.kBIGINT, 1, 0, 10**6)]
    table, names = gen_table(n, spec, frag_rows=1 << 21)
    rs, _ = gu.run_both(sqlmini.parse("SELECT key, SUM(v) FROM t GROUP BY key;", table, names), table, oracle_threads=8)
    assert rs.getQueryMemDesc().kernel == abi.KERNEL_PERFECT_GLOBAL
    # sparse variant: key = splitmix-like spread of the dense key => baseline hash, entry_count = NDV * 1.5
    sparse = abi.Table([(abi.kBIGINT, True), (abi.kBIGINT, True)])
    for f in table.fragments:
        k = f.host_cols[0].astype(np.uint64)
        k = ((k * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(1)).astype(np.int64)
        sparse.add_host_fragment([k, f.host_cols[1]])
    ndv = len(np.unique(np.concatenate([f.host_cols[0] for f in sparse.fragments])))
    rs, _ = gu.run_both(sqlmini.parse("SELECT key, SUM(v) FROM t GROUP BY key;", sparse, names), sparse,
                        entry_guess=int(ndv * 1.5), has_card=True, oracle_threads=1)
    assert rs.getQueryMemDesc().kernel == abi.KERNEL_BASELINE_GLOBAL and rs.rowCount() == ndv


def test_full_size_properties_2e8_rows():
    """Size-independent properties at a size the oracle cannot cover in seconds: with the filter open (100 %),
    SUM over groups of COUNT(*) == N; COUNT WHERE c0<k + COUNT WHERE c0>=k == N; group SUMs add up to the
    non-grouped SUM (which the oracle's generator formula can reproduce in closed form per column sample)."""
    import torch
    n = 200_000_000
    frag = 1 << 26
    bufs, dt = [], abi.Table([(abi.kBIGINT, True), (abi.kBIGINT, True), (abi.kINT, True)])
    for b in range(0, n, frag):
        m = min(frag, n - b)
        ptrs, stats = [], []
        for tag, (ty, lo, span) in enumerate([(abi.kBIGINT, 0, 10**6), (abi.kBIGINT, 0, 10**6), (abi.kINT, 0, 10**4)]):
            t = torch.empty(m * abi.SIZE_OF[ty], dtype=torch.uint8, device="cuda")
            executor.gen_column_device(t.data_ptr(), ty, SEED, tag, b, m, lo, span)
            bufs.append(t)
            ptrs.append(t.data_ptr())
            st = abi.ChunkStats()
            st.int_min, st.int_max = lo, lo + span - 1
            stats.append(st)
        dt.add_device_fragment(m, ptrs, stats)
    torch.cuda.synchronize()
    names = ["c0", "c1", "g"]
    ex = executor.Executor()

    def run(sql):
        return ex.executeWorkUnit(0, True, dt, sqlmini.parse(sql, dt, names), memory_level=abi.GPU_LEVEL).rows()
    grouped = run("SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 1000000 GROUP BY g;")
    assert len(grouped) == 10**4 and sum(r[2] for r in grouped) == n
    total = run("SELECT SUM(c1), COUNT(*) FROM t;")[0]
    assert total[1] == n and sum(r[1] for r in grouped) == total[0]
    lt = run("SELECT COUNT(*), SUM(c1) FROM t WHERE c0 < 500000;")[0]
    ge = run("SELECT COUNT(*), SUM(c1) FROM t WHERE c0 >= 500000;")[0]
    assert lt[0] + ge[0] == n and lt[1] + ge[1] == total[0]
    half = run("SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;")
    assert sum(r[2] for r in half) == lt[0] and sum(r[1] for r in half) == lt[1]
    # idempotence: the same query twice gives the identical buffer
    assert run("SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;") == half
    # a prefix the oracle CAN cover: first 2e6 rows of fragment 0 regenerated on the host
    m = 2_000_000
    host = abi.Table([(abi.kBIGINT, True), (abi.kBIGINT, True), (abi.kINT, True)])
    host.add_host_fragment([oracle_lib.gen_column(abi.kBIGINT, SEED, 0, 0, m, 0, 10**6), oracle_lib.gen_column(abi.kBIGINT, SEED, 1, 0, m, 0, 10**6),
                            oracle_lib.gen_column(abi.kINT, SEED, 2, 0, m, 0, 10**4)])
    pre = abi.Table(host.col_types)
    pre.add_device_fragment(m, dt.fragments[0].dev_ptrs, dt.fragments[0].stats)
    unit = sqlmini.parse("SELECT g, SUM(c1), COUNT(*) FROM t WHERE c0 < 500000 GROUP BY g;", host, names)
    got = ex.executeWorkUnit(0, True, pre, unit, memory_level=abi.GPU_LEVEL).rows()
    gu.rows_equal(got, oracle_lib.execute(unit, host, num_threads=8).rows())


def test_fragment_skipping_on_chunk_stats():
    """Executor::skipFragment (Execute.cpp:4776): fragments whose chunk min/max cannot satisfy a simple qual are not
    scanned (nor copied, for host tables); the answer is unchanged."""
    t = abi.Table([(abi.kINT, True), (abi.kBIGINT, True), (abi.kDOUBLE, True)])
    rng = np.random.default_rng(3)
    for f in range(6):   # disjoint key ranges per fragment: [100 f, 100 f + 99]
        n = 5000
        t.add_host_fragment([rng.integers(100 * f, 100 * f + 100, n).astype(np.int32), rng.integers(0, 1000, n),
                             rng.random(n) + f])
    names = ["a", "v", "d"]
    for sql, skipped in [("SELECT COUNT(*), SUM(v) FROM t WHERE a >= 200 AND a < 400;", 4),
                         ("SELECT a, COUNT(*) FROM t WHERE a = 350 GROUP BY a;", 5),
                         ("SELECT COUNT(*) FROM t WHERE d > 4.5;", 4),
                         ("SELECT COUNT(*) FROM t WHERE a > 1000;", 6),
                         ("SELECT COUNT(*) FROM t WHERE a < 250 OR a > 450;", 0)]:
        unit = sqlmini.parse(sql, t, names)
        for resident in (True, False):
            rs, _ = gu.run_both(unit, t, entry_guess=64, has_card=True, device_resident=resident)
            st = rs.stats()
            assert st["fragments_skipped"] == skipped and st["fragments_scanned"] == 6 - skipped, (sql, st)
            if not resident:
                assert (st["h2d_bytes"] == 0) == (skipped == 6)


def test_strided_generator_matches_oracle():
    import torch
    n = 50001
    want = oracle_lib.gen_column(abi.kBIGINT, SEED, 0, 777, n, 0, 10**7, stride=900_000_000_007)
    buf = torch.empty(n * 8, dtype=torch.uint8, device="cuda")
    executor.gen_column_device(buf.data_ptr(), abi.kBIGINT, SEED, 0, 777, n, 0, 10**7, stride=900_000_000_007)
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy().view(np.int64), want)


# ---- ENCODING FIXED chunks, deleted-rows column -----------------------------------------------
import enc_tables as et  # noqa: E402
import str_tables as stt  # noqa: E402


@pytest.mark.parametrize("n,frag_rows", [(1, 5), (999, 100), (60000, 8192)])
def test_fixed_encodings_and_deleted_rows(n, frag_rows):
    table = et.enc_table(n, seed=n, frag_rows=frag_rows)
    dev = gu.DeviceTable(table)
    for sql in et.ENC_QUERIES:
        unit = sqlmini.parse(sql, table, et.ENC_NAMES)
        gu.run_both(unit, table, entry_guess=120001, has_card=True, dev_table=dev)
    for sql in et.ENC_QUERIES[:4]:
        gu.run_both(sqlmini.parse(sql, table, et.ENC_NAMES), table, entry_guess=120001, has_card=True, device_resident=False)


def test_fully_deleted_fragment_is_skipped_and_flag_can_be_ignored():
    table = et.enc_table(4000, seed=11, frag_rows=1000, fully_deleted_fragment=2)
    unit = sqlmini.parse("SELECT k_i64_f32, COUNT(*), SUM(a_i64_f16) FROM e GROUP BY k_i64_f32;", table, et.ENC_NAMES)
    rs, _ = gu.run_both(unit, table)
    assert rs.stats()["fragments_skipped"] == 1          # isFragmentFullyDeleted
    # filter_on_deleted_column = false: all 4000 rows counted
    co = executor.compilation_options(filter_on_deleted_column=False)
    unit = sqlmini.parse("SELECT COUNT(*) FROM e;", table, et.ENC_NAMES)
    rs = executor.Executor().executeWorkUnit(0, True, table, unit, co=co)
    assert rs.rows() == [(4000,)]
    oracle_lib.lib().oracle_set_filter_on_deleted_column(0)
    try:
        assert oracle_lib.execute(unit, table).rows() == [(4000,)]
    finally:
        oracle_lib.lib().oracle_set_filter_on_deleted_column(1)


def test_inner_entry_b2q_launch_param_block():
    """The inner entry: the static-kernel replacement of the JIT'd multifrag_query_hoisted_literals call, driven with the
    reference's own 15-slot parameter block (enum KernelParam, QueryEngine/enums.h:64-79): COL_BUFFERS, NUM_FRAGMENTS,
    NUM_ROWS, GROUPBY_BUF (receives the finished reference-layout buffer on the DEVICE), ERROR_CODE."""
    import ctypes as C
    import torch
    table = random_table(30000, seed=77, frag_rows=7000)
    dev = gu.DeviceTable(table)
    L = executor.lib()
    for sql in ["SELECT nn32, COUNT(*), SUM(a64), MIN(a32), AVG(d) FROM r WHERE nn64 < 25 GROUP BY nn32;",
                "SELECT COUNT(*), SUM(big), MAX(a16) FROM r WHERE a8 <> 3;",
                "SELECT k64, SUM(a32) FROM r GROUP BY k64;"]:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        bt = dev.table.build(abi.GPU_LEVEL)
        co, eo = executor.compilation_options(), executor.execution_options()
        q = C.c_void_p()
        assert L.b2q_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(co), C.byref(eo), 0, 0, C.byref(q)) == 0
        plan = L.b2q_query_plan(q).contents
        nf, nc = len(dev.table.fragments), dev.table.num_cols
        # COL_BUFFERS: host array [frag] of host arrays [col] of device pointers (QueryExecutionContext.cpp:751-765)
        per_frag = [(C.c_void_p * nc)(*[p or None for p in f.dev_ptrs]) for f in dev.table.fragments]
        col_buffers = (C.POINTER(C.c_void_p) * nf)(*[C.cast(a, C.POINTER(C.c_void_p)) for a in per_frag])
        num_rows = (C.c_int64 * nf)(*[f.num_tuples for f in dev.table.fragments])
        num_frags, num_tables = C.c_uint32(nf), C.c_uint32(1)
        out = torch.zeros(max(int(plan.buffer_size), 8), dtype=torch.uint8, device="cuda")
        gb = torch.tensor([out.data_ptr()], dtype=torch.int64, device="cuda")      # int64_t** on the device
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        prm = abi.Params()
        prm.error_codes = err.data_ptr()
        prm.group_by_buffers = gb.data_ptr()
        prm.num_fragments, prm.num_tables = C.pointer(num_frags), C.pointer(num_tables)
        prm.col_buffers = col_buffers
        prm.num_rows = num_rows
        torch.cuda.synchronize()
        rc = L.b2q_launch(q, C.byref(prm), None)
        assert rc == 0, L.b2q_last_error_message()
        torch.cuda.synchronize()
        assert int(err.item()) == 0
        ref = oracle_lib.execute(unit, table, num_threads=4)
        got = out.cpu().numpy()[: int(plan.buffer_size)].view(np.int8)
        n = ref.entry_count()
        empty = np.array([bool(oracle_lib.lib().oracle_result_is_row_at_empty(ref.h, i)) for i in range(n)], dtype=bool)
        gu.buffers_equal(got, ref.buffer(), ref.plan, empty=empty)
        L.b2q_query_free(q)


# ---- dictionary-encoded string keys (int32 / uint8 / uint16 ids) and TIME-family columns ---------------------
@pytest.mark.parametrize("n,frag_rows", [(5, 2), (4000, 900), (150000, 40000)])
def test_dict_string_and_time_columns(n, frag_rows):
    table = stt.str_table(n, seed=12 + n, frag_rows=frag_rows)
    dev = gu.DeviceTable(table)
    for sql in stt.STR_QUERIES:
        unit = sqlmini.parse(sql, table, stt.STR_NAMES)
        if unit.unit.num_order_entries:
            continue     # ordered variants: test_gpu_order_by
        gu.run_both(unit, table, dev_table=dev)
        gu.run_both(unit, table, device_resident=False)
        if " GROUP BY dd" in sql or " GROUP BY dt" in sql:     # DATE keys are 8-byte keys: the columnar layout applies
            gu.run_both(unit, table, dev_table=dev, output_columnar=True)
    for sql in stt.STR_REJECTED:
        with pytest.raises(executor.UnsupportedOnThisPath):
            executor.Executor().executeWorkUnit(0, True, dev.table, sqlmini.parse(sql, table, stt.STR_NAMES), memory_level=abi.GPU_LEVEL)


def test_groupbytest_dictionary_key():
    """Tests/GroupByTest.cpp:73-152 (PerfectHashNoFallback) with its own key type: GROUP BY a dictionary string."""
    t = abi.Table([(abi.kINT, True), (abi.kTEXT, False)])
    t.add_host_fragment([np.array([1, 2], dtype=np.int32), np.array([0, 1], dtype=np.int32)])
    rs, _ = gu.run_both(sqlmini.parse("SELECT COUNT(*) FROM t WHERE x = 1 GROUP BY str;", t, ["x", "str"]), t)
    assert rs.rowCount() == 1 and rs.rows() == [(1,)] and rs.getQueryMemDesc().query_desc_type == abi.GroupByPerfectHash
    rs, _ = gu.run_both(sqlmini.parse("SELECT str, COUNT(*) FROM t GROUP BY str;", t, ["x", "str"]), t)
    assert sorted(rs.rows()) == [(0, 1), (1, 1)]


def test_groupbytest_baseline_fallback():
    """Tests/GroupByTest.cpp:173-262 (BaselineFallbackTest) through the CUDA path: first CardinalityEstimationRequired, then — with
    has_cardinality_estimation and max_groups_buffer_entry_guess = 1 — one row whose value is 1; radix passes and per-row probe."""
    from test_dict_strings import high_cardinality_str_table
    t = high_cardinality_str_table()
    unit = sqlmini.parse("SELECT COUNT(*) FROM t WHERE x = 1 GROUP BY str;", t, ["x", "str"])
    with pytest.raises(executor.CardinalityEstimationRequired):
        executor.Executor().executeWorkUnit(1, True, t, unit, has_cardinality_estimation=False)
    for force in (0, abi.KERNEL_BASELINE_PROBE):
        rs, _ = gu.run_both(unit, t, entry_guess=1, has_card=True, force_kernel=force)
        assert rs.rowCount() == 1 and rs.rows() == [(1,)] and rs.getQueryMemDesc().query_desc_type == abi.GroupByBaselineHash
