"""The reference's own golden query strings, verbatim: every `c("SELECT ... FROM test ...", dt)` of Tests/ExecuteTest.cpp that this
path accepts (tests/golden/executetest_harvest.json, made by tools/harvest_executetest.py where the reference tree exists — ~120 of the
~600 single-table strings, over the numeric, dictionary-string and FIXED columns of the golden table (tests/ref_full_table.py); the rest use
expressions, CASE, HAVING, string functions, none-encoded strings ... outside the path).  Expected rows come from
SQLite over the same golden table, as the reference's SQLiteComparator computes them (ExecuteTest.cpp:383-520); the product's planner
must decide what the oracle decides, and the product's host side (lowered program emulated on the host + read-out) must return the
same rows."""
import json
import os

import pytest

import oracle_lib
import ref_full_table as ft
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_filter_lowering import emu, run_program  # noqa: F401  (emu is a fixture)
from test_oracle_fuzz import known_reference_quirk

HERE = os.path.dirname(os.path.abspath(__file__))
HARVEST = json.load(open(os.path.join(HERE, "golden", "executetest_harvest.json")))
QUERIES = [q["sql"] for q in HARVEST["queries"]]


@pytest.fixture(scope="module")
def golden():
    rows = ft.full_rows()
    return ft.make_table(rows), ft.make_sqlite(rows)


def test_harvest_is_what_the_script_reports():
    assert HARVEST["stats"]["kept"] == len(QUERIES) >= 60 and len(set(QUERIES)) == len(QUERIES)


@pytest.mark.parametrize("sql", QUERIES)
def test_reference_query_verbatim(golden, emu, sql):  # noqa: F811
    table, con = golden
    unit = sqlmini.parse(sql, table, ft.FULL_NAMES, dicts=ft.DICTS)
    res = oracle_lib.execute(unit, table, entry_guess=48, has_card=True, num_threads=2)
    got = executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=48, has_cardinality_estimation=True).as_dict()
    assert got == res.plan.as_dict()
    if known_reference_quirk(unit, res.plan):
        pytest.skip("keyless marker on a nullable MIN / MAX / SUM: the reference itself departs from SQLite (test_reference_quirks)")
    ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
    if unit.unit.num_order_entries:
        from test_order_by import assert_ordered_rows_match
        import order_queries as oq
        ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "test")).fetchall()]
        assert_ordered_rows_match(ft.translate_strings(res.rows(), res.plan), ref)
    else:
        fp_abs = rt.float_sum_atol(len(rt.test_rows())) if any(t.is_agg and t.agg_kind in (abi.kSUM, abi.kAVG) and t.agg_arg_type.type == abi.kFLOAT
                                                              for t in res.plan.targets[: res.plan.num_targets]) else 0.0
        rt.assert_rows_match(ft.translate_strings(res.rows(), res.plan), ref, fp_abs=fp_abs)
        # the product without its CUDA kernels: b2q_plan -> the lowered program run on the host -> b2q_rs_create_from_storage -> rows
        if res.plan.query_desc_type in (abi.GroupByPerfectHash, abi.NonGroupedAggregate) and not unit.unit.has_limit and not unit.unit.offset:
            rc, buf = run_program(emu, unit, table, entry_guess=48, has_card=True)
            assert rc == 0, (sql, rc)
            rs = executor.Executor().resultSetFromStorage(buf, unit, table, max_groups_buffer_entry_guess=48, has_cardinality_estimation=True)
            rt.assert_rows_match(ft.translate_strings(rs.rows(), res.plan), ref, fp_abs=fp_abs)
