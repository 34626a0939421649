"""Randomised queries, oracle vs SQLite (CPU only): the same generators that drive the GPU fuzz tests pin the ORACLE
here — filter trees with NOT / IS NULL / IN, every aggregate over every column type, single / composite / baseline keys,
INNER and LEFT star joins — against an independent SQL engine, the way Tests/ExecuteTest.cpp uses SQLite."""
import random
import re

import pytest

import join_tables as jt
import oracle_lib
import order_queries as oq
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_gpu_fuzz import rand_join_query, rand_query
from test_gpu_parity import RAND_COLS, RAND_NAMES, random_table


def known_reference_quirk(unit, plan) -> bool:
    """Shapes where the reference itself departs from SQL (see test_oracle_golden.test_reference_quirks): a keyless
    layout whose marker slot is a MIN/MAX/SUM that can legitimately stay at its init value (all-NULL group)."""
    if not plan.keyless_hash or plan.idx_target_as_key < 0:
        return False
    for t in plan.targets[: plan.num_targets]:
        if t.first_slot == plan.idx_target_as_key or (t.agg_kind == abi.kAVG and t.first_slot + 1 == plan.idx_target_as_key):
            return t.is_agg and t.agg_kind in (abi.kMIN, abi.kMAX, abi.kSUM) and not t.agg_arg_type.notnull
    return False


def float_sum_tol(sql: str, table) -> dict:
    """Tolerances for a query that sums a FLOAT column (float accumulation, ref_tables.float_sum_atol)."""
    if re.search(r"(SUM|AVG)\((f32|fnn)\)", sql):
        return {"fp_abs": rt.float_sum_atol(sum(f.num_tuples for f in table.fragments))}
    return {}


def sqlite_overflows(sql: str) -> bool:
    return "SUM(big)" in sql or "AVG(big)" in sql or "SUM(d.big)" in sql or "AVG(d.big)" in sql   # SQLite raises on int64 overflow, HeavyDB wraps


@pytest.mark.parametrize("seed", range(4))
def test_single_table_queries(seed):
    rng = random.Random(9000 + seed)
    table = random_table([40, 1500, 4000, 4000][seed], seed=300 + seed, frag_rows=[7, 400, 4000, 900][seed])
    con = rt.make_sqlite(oq.rows_of(table, RAND_COLS), RAND_COLS, "r")
    checked = planned = 0
    for i in range(120):
        sql = rand_query(rng, multi_key=(i % 3 == 0))
        if sqlite_overflows(sql):
            continue
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            res = oracle_lib.execute(unit, table, entry_guess=6000, has_card=True, num_threads=2)
        except oracle_lib.OracleError as e:
            assert e.code == abi.ERR_UNSUPPORTED, sql
            continue
        # the product's planner accepts what the oracle accepts and decides the same (the device filter program holds
        # 16 leaves on a 4-deep mask stack; only a filter past that may be refused)
        try:
            got = executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=6000, has_cardinality_estimation=True).as_dict()
            assert got == res.plan.as_dict(), sql
            planned += 1
        except executor.UnsupportedOnThisPath as e:
            assert "filter" in str(e), sql
        if known_reference_quirk(unit, res.plan):
            continue
        ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
        try:
            rt.assert_rows_match(res.rows(), ref, **float_sum_tol(sql, table))
        except AssertionError as e:
            raise AssertionError(f"query: {sql}\n{e}") from e
        checked += 1
    assert checked >= 60 and planned >= checked


@pytest.mark.parametrize("seed", range(3))
def test_join_queries(seed):
    rng = random.Random(7000 + seed)
    fact = jt.fact_table([60, 3000, 6000][seed], seed=80 + seed, frag_rows=[25, 800, 6000][seed])
    dim = jt.dim_table(seed=17 + seed)
    con = rt.make_sqlite(jt.logical_rows(fact, jt.FACT_COLS), jt.FACT_COLS, "t")
    decl = ", ".join(f"{n} {'double' if t == abi.kDOUBLE else 'bigint'}" for n, t, _ in jt.DIM_COLS)
    con.execute(f"CREATE TABLE d({decl})")
    con.executemany(f"INSERT INTO d VALUES({','.join('?' * len(jt.DIM_COLS))})", jt.logical_rows(dim, jt.DIM_COLS))
    checked = 0
    for _ in range(100):
        sql = rand_join_query(rng)
        if sqlite_overflows(sql):
            continue
        unit = sqlmini.parse(sql, fact, jt.FACT_NAMES, inner=(dim, jt.DIM_NAMES))
        try:
            res = oracle_lib.execute(unit, fact, entry_guess=6000, has_card=True, num_threads=2)
        except oracle_lib.OracleError as e:
            assert e.code == abi.ERR_UNSUPPORTED, sql
            continue
        if known_reference_quirk(unit, res.plan):
            continue
        ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
        try:
            rt.assert_rows_match(res.rows(), ref)
        except AssertionError as e:
            raise AssertionError(f"query: {sql}\n{e}") from e
        checked += 1
    assert checked >= 50


def with_random_order(rng, sql: str):
    """Append ORDER BY over random targets followed by every GROUP BY key (a total order), plus LIMIT / OFFSET.
    Returns None when the keys are not all in the target list (positions could not name them)."""
    body = sql.rstrip(";")
    up = body.upper()
    if " GROUP BY " not in up:
        return None
    targets = [t.strip() for t in body[len("SELECT "):up.index(" FROM ")].split(",")]
    keys = [k.strip() for k in body[up.index(" GROUP BY ") + 10:].split(",")]
    if any(k not in targets for k in keys):
        return None
    items = []
    for pos in rng.sample(range(1, len(targets) + 1), rng.randint(1, min(3, len(targets)))):
        if targets[pos - 1] in keys:
            continue
        items.append(f"{pos} {rng.choice(['ASC', 'DESC'])} NULLS {rng.choice(['FIRST', 'LAST'])}")
    for k in keys:
        items.append(f"{targets.index(k) + 1} {rng.choice(['ASC', 'DESC'])} NULLS {rng.choice(['FIRST', 'LAST'])}")
    tail = ""
    if rng.random() < 0.6:
        tail = f" LIMIT {rng.randint(1, 40)}"
        if rng.random() < 0.4:
            tail += f" OFFSET {rng.randint(0, 10)}"
    return body + " ORDER BY " + ", ".join(items) + tail + ";"


@pytest.mark.parametrize("seed", range(3))
def test_ordered_queries(seed):
    """ResultSet::sort / dropFirstN / keepFirstN restatement on random ORDER BY lists, compared row by row."""
    from test_order_by import assert_ordered_rows_match
    rng = random.Random(5000 + seed)
    table = random_table([300, 3000, 3000][seed], seed=500 + seed, frag_rows=[64, 700, 3000][seed])
    con = rt.make_sqlite(oq.rows_of(table, RAND_COLS), RAND_COLS, "r")
    checked = 0
    for i in range(150):
        sql = with_random_order(rng, rand_query(rng, multi_key=(i % 4 == 0)))
        if sql is None or sqlite_overflows(sql):
            continue
        # AVG / SUM of doubles as an order key could tie-break differently on the last ulp: keep exact key types
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            res = oracle_lib.execute(unit, table, entry_guess=6000, has_card=True, num_threads=2)
        except oracle_lib.OracleError as e:
            assert e.code == abi.ERR_UNSUPPORTED, sql
            continue
        if known_reference_quirk(unit, res.plan):
            continue
        ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "r")).fetchall()]
        try:
            assert_ordered_rows_match(res.rows(), ref, **float_sum_tol(sql, table))
        except AssertionError as e:
            raise AssertionError(f"query: {sql}\n{e}") from e
        checked += 1
    assert checked >= 40
