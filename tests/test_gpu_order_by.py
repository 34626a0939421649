"""ORDER BY / LIMIT / OFFSET on the device (sort.cu: compaction + stable radix sort + gather) against the oracle's
restatement of ResultSet::sort / dropFirstN / keepFirstN — rows in order, and the compact buffer row for row."""
import numpy as np
import pytest

import gpu_util as gu
import oracle_lib
import order_queries as oq
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_gpu_parity import RAND_NAMES, random_table

pytestmark = pytest.mark.gpu


def window(unit, n_sorted):
    """[first, first + count) of the oracle's permutation that the product materialises."""
    u = unit.unit
    if u.has_limit and u.limit == 0:      # RelSort::isEmptyResult(): an empty result, not "no limit"
        return 0, 0
    first = min(u.offset, n_sorted)
    count = n_sorted - first
    if u.has_limit and u.limit:
        count = min(count, u.limit)
    return first, count


def run_sorted(unit, table, dev, output_columnar=False, **kw):
    ex = executor.Executor()
    eo = executor.execution_options(output_columnar_hint=output_columnar)
    rs = ex.executeWorkUnit(kw.get("entry_guess", 0), True, dev.table, unit, eo=eo,
                            has_cardinality_estimation=kw.get("has_card", False), memory_level=abi.GPU_LEVEL)
    ref = oracle_lib.execute(unit, table, entry_guess=kw.get("entry_guess", 0), has_card=kw.get("has_card", False),
                             num_threads=4, output_columnar=output_columnar)
    gu.rows_equal_ordered(rs.rows(), ref.rows())
    assert rs.rowCount() == ref.row_count()
    # the compact buffer: entry i of the product == entry perm[first + i] of the oracle's (full) buffer
    gp, op = rs.getQueryMemDesc(), ref.plan
    if unit.unit.num_order_entries or unit.unit.has_limit or unit.unit.offset:
        perm = ref.permutation() if unit.unit.num_order_entries else \
            [e for e in range(op.entry_count) if not oracle_lib.lib().oracle_result_is_row_at_empty(ref.h, e)]
        first, count = window(unit, len(perm))
        assert rs.entryCount() == count == gp.entry_count
        gu.compact_buffer_equal(rs.getStorageBuffer(), gp, ref.buffer(), op, perm[first:first + count])
    return rs, ref


def test_golden_order_by():
    table = rt.make_table(rt.test_rows())
    dev = gu.DeviceTable(table)
    for sql in oq.GOLDEN_ORDER_QUERIES:
        unit = sqlmini.parse(sql, table, rt.TEST_NAMES)
        try:
            run_sorted(unit, table, dev, entry_guess=64, has_card=True)
        except Exception as e:
            raise AssertionError(f"query: {sql}\n{e}") from e


@pytest.mark.parametrize("n,frag_rows", [(3000, 700), (200000, 65536)])
def test_random_order_by(n, frag_rows):
    table = random_table(n, seed=31, frag_rows=frag_rows)
    dev = gu.DeviceTable(table)
    for sql in oq.RAND_ORDER_QUERIES + [oq.OFFSET_WITHOUT_LIMIT_QUIRK]:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            run_sorted(unit, table, dev, entry_guess=3001, has_card=True)
        except Exception as e:
            raise AssertionError(f"query: {sql}\n{e}") from e


def test_order_by_columnar_output():
    table = random_table(50000, seed=8, frag_rows=20000)
    dev = gu.DeviceTable(table)
    ran = 0
    for sql in oq.RAND_ORDER_QUERIES:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            oracle_lib.plan(unit, table, entry_guess=3001, has_card=True, output_columnar=True)
        except oracle_lib.OracleError:
            continue
        rs, _ = run_sorted(unit, table, dev, output_columnar=True, entry_guess=3001, has_card=True)
        assert rs.getQueryMemDesc().output_columnar == 1
        ran += 1
    assert ran >= 5


def test_order_by_on_time_and_dict_tables():
    import str_tables as stt
    table = stt.str_table(20000, seed=3, frag_rows=6000)
    dev = gu.DeviceTable(table)
    for sql in [q for q in stt.STR_QUERIES if " ORDER BY " in q] + [
            "SELECT ts, COUNT(*), MAX(dt) FROM s GROUP BY ts ORDER BY 1 DESC NULLS LAST LIMIT 20;",
            "SELECT s8, COUNT(str) FROM s GROUP BY s8 ORDER BY 2 DESC, 1 LIMIT 7;" if False else
            "SELECT x, COUNT(str), MIN(ts) FROM s GROUP BY x ORDER BY 3 ASC NULLS FIRST, 1;"]:
        unit = sqlmini.parse(sql, table, stt.STR_NAMES)
        run_sorted(unit, table, dev)


def test_result_set_sort_api():
    """b2q_rs_sort / drop_first_n / keep_first_n on a finished (unsorted) result set == sort_info in the unit."""
    table = random_table(40000, seed=5, frag_rows=9000)
    dev = gu.DeviceTable(table)
    sql = "SELECT nn32, SUM(a32), AVG(a32) FROM r GROUP BY nn32 ORDER BY 3 DESC NULLS LAST, 1 LIMIT 25 OFFSET 10;"
    ex = executor.Executor()
    want = ex.executeWorkUnit(0, True, dev.table, sqlmini.parse(sql, table, RAND_NAMES), memory_level=abi.GPU_LEVEL).rows()
    rs = ex.executeWorkUnit(0, True, dev.table, sqlmini.parse(sql[:sql.index(" ORDER BY")] + ";", table, RAND_NAMES),
                            memory_level=abi.GPU_LEVEL)
    n_all = rs.rowCount()
    rs.sort([(3, True, False), (1, False, False)], top_n=35)
    assert rs.entryCount() == 35
    rs.dropFirstN(10)
    rs.keepFirstN(25)
    assert rs.rows() == want and rs.rowCount() == 25
    rs.dropFirstN(0)
    rs.keepFirstN(0)
    rs.sort([(1, False, False)])        # full sort by the key
    rows = rs.rows()
    assert len(rows) == n_all and [r[0] for r in rows] == sorted(r[0] for r in rows)


def test_topk_prefilter_with_nulls():
    """> 65536 groups and a small LIMIT: the top-k pre-filter (sort.cu, topk_bucket) runs; NULL ranks, DESC, ties on the
    primary key and a second order entry must all survive it."""
    table = random_table(400000, seed=77, frag_rows=100000)
    dev = gu.DeviceTable(table)
    for sql in ["SELECT a64, COUNT(*), SUM(a32) FROM r GROUP BY a64 ORDER BY 3 DESC NULLS LAST, 1 LIMIT 20;",
                "SELECT a64, COUNT(*), SUM(a32) FROM r GROUP BY a64 ORDER BY 3 ASC NULLS FIRST, 1 DESC NULLS LAST LIMIT 50 OFFSET 5;",
                "SELECT a64, COUNT(*), MIN(a8) FROM r GROUP BY a64 ORDER BY 2 DESC, 3 ASC NULLS LAST, 1 LIMIT 100;",     # heavy ties on COUNT
                "SELECT a64, AVG(d), COUNT(*) FROM r GROUP BY a64 ORDER BY 2 DESC NULLS LAST, 1 LIMIT 7;"]:
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        rs, _ = run_sorted(unit, table, dev, entry_guess=1_000_000, has_card=True)
        assert rs.rowCount() > 0


def test_large_sort_c4_like():
    """1e6 dense int64 keys, ORDER BY SUM DESC LIMIT 10 / full sort: multi-block compaction, all radix passes."""
    n, keys = 4_000_000, 1_000_000
    k = oracle_lib.gen_column(abi.kBIGINT, 0x5EED, 1, 0, n, 0, keys)
    v = oracle_lib.gen_column(abi.kBIGINT, 0x5EED, 2, 0, n, 0, 1_000_000)
    table = abi.Table([(abi.kBIGINT, True), (abi.kBIGINT, True)])
    table.add_host_fragment([k, v])
    dev = gu.DeviceTable(table)
    for sql in ["SELECT key, SUM(v) FROM t GROUP BY key ORDER BY 2 DESC, 1 LIMIT 10;",
                "SELECT key, SUM(v), COUNT(*) FROM t GROUP BY key ORDER BY 3, 2 DESC, 1 LIMIT 1000 OFFSET 999000;"]:
        unit = sqlmini.parse(sql, table, ["key", "v"])
        rs, ref = run_sorted(unit, table, dev)
        assert rs.stats()["sort_us"] > 0
