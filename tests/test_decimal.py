"""DECIMAL / NUMERIC columns (value x 10^scale as int64, or 32 / 16-bit FIXED chunks): integers on the whole path, the
scale applied at read-out only (makeTargetValue ResultSetIteration.cpp:2193-2210, pair_to_double
ResultSetBufferAccessors.h:197-227).  The oracle against SQLite on the reference's golden `dd` / `dd_notnull` columns
(ExecuteTest.cpp:1971-1986, :2823, :12022) and on a table with NULLs; the product's planner against the oracle's; the lowered
program, read on the host, against the oracle's buffer."""
import pytest

import dec_tables as dt
import oracle_lib
import order_queries as oq
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_filter_lowering import assert_buffers_match, emu, run_program  # noqa: F401  (emu is a fixture)
from test_order_by import assert_ordered_rows_match


@pytest.fixture(scope="module")
def golden():
    rows = dt.golden_rows()
    return dt.make_table(rows), dt.make_sqlite(rows)


@pytest.fixture(scope="module")
def mixed():
    rows = dt.mixed_rows()
    return dt.make_table(rows, fragment_size=170), dt.make_sqlite(rows)


def _check(env, sql):
    table, con = env
    unit = sqlmini.parse(sql, table, dt.DEC_NAMES)
    res = oracle_lib.execute(unit, table, num_threads=2)
    ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "test")).fetchall()]
    if unit.unit.num_order_entries:
        assert_ordered_rows_match(res.rows(), ref)
    else:
        rt.assert_rows_match(res.rows(), ref)
    assert executor.Executor().plan(unit, table).as_dict() == oracle_lib.plan(unit, table).as_dict()
    return unit, res


@pytest.mark.parametrize("sql", dt.GOLDEN_QUERIES)
def test_golden_dd_vs_sqlite(golden, sql):
    _check(golden, sql)


@pytest.mark.parametrize("sql", dt.GOLDEN_QUERIES + dt.MORE_QUERIES)
def test_nullable_and_fixed_width_decimals_vs_sqlite(mixed, sql):
    _check(mixed, sql)


def test_known_answers(golden):
    """What follows from the three INSERT templates (10 x 111.1, 5 x 222.2, 5 x 333.3)."""
    table, _ = golden
    def rows(sql, **kw):
        return oracle_lib.execute(sqlmini.parse(sql, table, dt.DEC_NAMES), table).rows(**kw)
    assert rows("SELECT COUNT(*) FROM test WHERE dd > 111.0;") == [(20,)]
    assert rows("SELECT COUNT(*) FROM test WHERE dd > 222.0;") == [(10,)]
    assert rows("SELECT COUNT(*) FROM test WHERE dd > 333.0;") == [(5,)]
    assert rows("SELECT COUNT(*) FROM test WHERE dd > 333.3;") == [(0,)]
    # decimal_to_double = false hands out the scaled integers (ResultSetIteration.cpp:2209)
    assert rows("SELECT MIN(dd), MAX(dd), SUM(dd) FROM test;", decimal_to_double=False) == [(11110, 33330, 10 * 11110 + 5 * 22220 + 5 * 33330)]
    assert rows("SELECT MIN(dd), MAX(dd), SUM(dd) FROM test;") == [(111.1, 333.3, 3888.5)]
    (avg,), = rows("SELECT AVG(dd) FROM test;")
    assert avg == 388850 / (20 * 100.0)          # ONE division by count x 10^scale (ResultSetBufferAccessors.h:222-225)
    assert rows("SELECT dd, COUNT(*) FROM test GROUP BY dd;", decimal_to_double=False) == [(11110, 10), (22220, 5), (33330, 5)]


def test_types_carry_the_scale(golden):
    table, _ = golden
    unit = sqlmini.parse("SELECT dd, SUM(dd), AVG(dd), MIN(p), COUNT(dd) FROM test GROUP BY dd;", table, dt.DEC_NAMES)
    res = oracle_lib.execute(unit, table)
    assert [res.col_type(i) for i in range(5)] == [(abi.kDECIMAL, 0, 2), (abi.kDECIMAL, 0, 2), (abi.kDOUBLE, 0, 0), (abi.kDECIMAL, 0, 3), (abi.kINT, 0, 0)]
    plan = executor.Executor().plan(unit, table)
    assert plan.as_dict() == res.plan.as_dict()
    # AVG keeps the DECIMAL target type (SQLTypeInfo::is_integer() is false for it, TargetInfo.cpp:57-67)
    assert plan.targets[2].sql_type.type == abi.kDECIMAL and plan.targets[2].sql_type.scale == 2


def test_refused_on_both_sides(golden):
    """A DECIMAL against a value of another type or scale reaches the reference's executor under a CAST node."""
    table, _ = golden
    for build in (lambda b: b.cmp(2, abi.kGT, 111),                                   # integer constant
                  lambda b: b.cmp(2, abi.kGT, 111.0),                                 # DOUBLE constant
                  lambda b: b.cmp(2, abi.kGT, 111000, abi.kDECIMAL, scale=3),         # another scale
                  lambda b: b.binop(abi.kEQ, b.col(2), b.col(4)),                     # DECIMAL(10,2) = DECIMAL(7,3)
                  lambda b: b.cmp(0, abi.kGT, 700, abi.kDECIMAL, scale=2)):           # INT column against a DECIMAL constant
        b = abi.UnitBuilder(table)
        b.add_qual(build(b))
        b.target(b.agg(abi.kCOUNT))
        unit = b.build()
        with pytest.raises(oracle_lib.OracleError):
            oracle_lib.execute(unit, table)
        with pytest.raises(executor.UnsupportedOnThisPath):
            executor.Executor().plan(unit, table)
    with pytest.raises(ValueError):
        sqlmini.parse("SELECT COUNT(*) FROM test WHERE dd > 1.234;", table, dt.DEC_NAMES)


def test_lowered_program_reproduces_the_oracles_buffer(emu, mixed):  # noqa: F811
    table, _ = mixed
    ran = 0
    for sql in dt.GOLDEN_QUERIES + dt.MORE_QUERIES:
        if " ORDER BY " in sql:
            sql = sql[:sql.index(" ORDER BY ")] + ";"
        unit = sqlmini.parse(sql, table, dt.DEC_NAMES)
        for columnar in (False, True):
            try:
                res = oracle_lib.execute(unit, table, output_columnar=columnar)
            except oracle_lib.OracleError:
                continue
            if res.plan.query_desc_type not in (abi.GroupByPerfectHash, abi.NonGroupedAggregate):
                continue
            rc, got = run_program(emu, unit, table, output_columnar=columnar)
            assert rc == 0, (sql, rc)
            assert_buffers_match(got, res.buffer(), sql)
            ran += 1
    assert ran >= 30


def _rand_decimal_query(rng):
    """Random filter tree + aggregates + keys over the DECIMAL table; literals are written at the column's scale or coarser."""
    lits = {"x": lambda: str(rng.randint(0, 9)), "y": lambda: str(rng.randint(39, 45)),
            "dd": lambda: f"{rng.randint(-5200, 5200) / 100:.2f}", "dd_notnull": lambda: f"{rng.randint(0, 40) * 25 / 100:.2f}",
            "p": lambda: rng.choice([f"{rng.randint(-9999999, 9999999) / 1000:.3f}", str(rng.randint(-9999, 9999))]),
            "q": lambda: f"{rng.randint(-9999, 9999) / 100:.2f}"}
    def leaf():
        c = rng.choice(list(lits))
        k = rng.random()
        if k < 0.12 and c not in ("x", "dd_notnull"):
            return f"{c} IS {'NOT ' if rng.random() < 0.5 else ''}NULL"
        if k < 0.22:
            lo, hi = sorted([float(lits[c]()), float(lits[c]())])
            fmt = {"x": "%d", "y": "%d", "p": "%.3f"}.get(c, "%.2f")
            return f"{c} BETWEEN {fmt % lo} AND {fmt % hi}"
        if k < 0.3:
            return f"{c} IN ({', '.join(lits[c]() for _ in range(rng.randint(2, 4)))})"
        return f"{c} {rng.choice(['=', '<>', '<', '>', '<=', '>='])} {lits[c]()}"
    def tree(depth):
        if depth == 0 or rng.random() < 0.4:
            return leaf()
        return f"({tree(depth - 1)} {rng.choice(['AND', 'OR'])} {tree(depth - 1)})"
    aggs = ["COUNT(*)"]
    for _ in range(rng.randint(1, 4)):
        kind = rng.choice(["COUNT", "SUM", "MIN", "MAX", "AVG"])
        col = "dd_notnull" if kind in ("SUM", "AVG") else rng.choice(["dd", "dd_notnull", "p", "q", "y"])   # exact in binary: SQLite's REAL sums stay comparable
        aggs.append(f"{kind}({col})")
    keys = rng.sample(["x", "y", "dd_notnull", "q", "dd"], rng.choice([0, 1, 1, 2]))
    where = f" WHERE {tree(2)}" if rng.random() < 0.8 else ""
    group = f" GROUP BY {', '.join(keys)}" if keys else ""
    return f"SELECT {', '.join(keys + aggs)} FROM test{where}{group};"


@pytest.mark.parametrize("seed", range(3))
def test_random_decimal_queries(emu, mixed, seed):  # noqa: F811
    """Oracle vs SQLite, the product's planner vs the oracle's, and the lowered program read on the host vs the oracle's buffer."""
    import random
    from test_oracle_fuzz import known_reference_quirk
    table, con = mixed
    rng = random.Random(4200 + seed)
    checked = emulated = 0
    for _ in range(80):
        sql = _rand_decimal_query(rng)
        unit = sqlmini.parse(sql, table, dt.DEC_NAMES)
        try:
            res = oracle_lib.execute(unit, table, entry_guess=6000, has_card=True, num_threads=2)
        except oracle_lib.OracleError as e:
            assert e.code == abi.ERR_UNSUPPORTED, sql
            continue
        try:
            got = executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=6000, has_cardinality_estimation=True).as_dict()
            assert got == res.plan.as_dict(), sql
        except executor.UnsupportedOnThisPath as e:
            assert "filter" in str(e), sql
            continue
        if res.plan.query_desc_type in (abi.GroupByPerfectHash, abi.NonGroupedAggregate):
            rc, buf = run_program(emu, unit, table, entry_guess=6000, has_card=True)
            assert rc == 0, (sql, rc)
            assert_buffers_match(buf, res.buffer(), sql)
            emulated += 1
        if known_reference_quirk(unit, res.plan):
            continue
        ref = [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()]
        try:
            rt.assert_rows_match(res.rows(), ref)
        except AssertionError as e:
            raise AssertionError(f"query: {sql}\n{e}") from e
        checked += 1
    assert checked >= 40 and emulated >= 40


def test_decimal_attribute_of_a_joined_dimension(emu):  # noqa: F811
    """Star join whose dimension carries a DECIMAL(7, 2) price stored as FIXED(32): filter, key and aggregate argument read
    through the join index — oracle vs SQLite, product planner vs oracle, lowered join program vs the oracle's buffer (INNER and LEFT)."""
    import sqlite3
    fact, dim, price, fk_all, v_all = dt.star_join()
    con = sqlite3.connect(":memory:")
    con.execute("CREATE TABLE t(fk bigint, v bigint)")
    con.execute("CREATE TABLE d(id bigint, price double)")
    con.executemany("INSERT INTO t VALUES(?, ?)", [(None if a == abi.NULL_INT else a, b) for a, b in zip(fk_all, v_all)])
    con.executemany("INSERT INTO d VALUES(?, ?)", [(i, None if p == -2**31 else p / 100) for i, p in enumerate(price.tolist())])
    for sql in dt.JOIN_QUERIES:
        unit = sqlmini.parse(sql, fact, dt.FACT_NAMES, inner=(dim, dt.DIM_NAMES))
        res = oracle_lib.execute(unit, fact, num_threads=2)
        rt.assert_rows_match(res.rows(), [tuple(r) for r in con.execute(sql.rstrip(";")).fetchall()])
        assert executor.Executor().plan(unit, fact).as_dict() == res.plan.as_dict(), sql
        # the lowered join program, read on the host over the denormalised rows, gives the oracle's buffer
        from test_filter_lowering import run_join_program
        rc, got = run_join_program(emu, unit, fact, dim, left=" LEFT JOIN " in sql, entry_guess=0)
        assert rc == 0, (sql, rc)
        assert_buffers_match(got, res.buffer(), sql)
