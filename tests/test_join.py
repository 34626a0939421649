"""One-level INNER hash join with a one-to-one perfect table — CPU side: the oracle's restatement of
PerfectJoinHashTable (range, fill_hash_join_buff) + hash_join_idx[_nullable] in the row loop against SQLite, planner
parity, and what is refused (one-to-many, sparse ranges, outer joins)."""
import numpy as np
import pytest

import join_tables as jt
import oracle_lib
import order_queries as oq
import ref_tables as rt
import sqlmini
from heavydb_b200 import abi, executor
from test_order_by import assert_ordered_rows_match


@pytest.fixture(scope="module")
def env():
    fact = jt.fact_table(6000, seed=3, frag_rows=1700)
    dim = jt.dim_table()
    con = rt.make_sqlite(jt.logical_rows(fact, jt.FACT_COLS), jt.FACT_COLS, "t")
    decl = ", ".join(f"{n} {'double' if t == abi.kDOUBLE else 'bigint'}" for n, t, _ in jt.DIM_COLS)
    con.execute(f"CREATE TABLE d({decl})")
    con.executemany(f"INSERT INTO d VALUES({','.join('?' * len(jt.DIM_COLS))})", jt.logical_rows(dim, jt.DIM_COLS))
    return fact, dim, con


def parse(sql, fact, dim):
    return sqlmini.parse(sql, fact, jt.FACT_NAMES, inner=(dim, jt.DIM_NAMES))


@pytest.mark.parametrize("sql", jt.JOIN_QUERIES + jt.LEFT_JOIN_QUERIES)
def test_oracle_join_vs_sqlite(env, sql):
    fact, dim, con = env
    unit = parse(sql, fact, dim)
    res = oracle_lib.execute(unit, fact, entry_guess=4000, has_card=True, num_threads=3)
    ref = [tuple(r) for r in con.execute(oq.sqlite_sql(sql, unit, "t")).fetchall()]
    if unit.unit.num_order_entries:
        assert_ordered_rows_match(res.rows(), ref)
    else:
        rt.assert_rows_match(res.rows(), ref)
    want = oracle_lib.plan(unit, fact, entry_guess=4000, has_card=True).as_dict()
    got = executor.Executor().plan(unit, fact, max_groups_buffer_entry_guess=4000, has_cardinality_estimation=True).as_dict()
    assert got == want
    assert got["join_entry_count"] > 0 and got["join_outer_col"] >= 0


def test_join_table_known_answers(env):
    """The table the reference would build: hash_entry_count = max - min + 1 over the INNER key's range, int32 slots."""
    fact, dim, _ = env
    p = oracle_lib.plan(parse("SELECT COUNT(*) FROM t JOIN d ON t.fk32 = d.id32;", fact, dim), fact)
    assert (p.join_min_key, p.join_max_key, p.join_entry_count) == (3, jt.DIM_ROWS + 2, jt.DIM_ROWS)
    assert (p.join_outer_col, p.join_inner_col) == (0, 0)
    p = oracle_lib.plan(parse("SELECT COUNT(*) FROM t JOIN d ON t.fk64 = d.id64;", fact, dim), fact)
    assert p.join_entry_count == p.join_max_key - p.join_min_key + 1 and p.join_entry_count <= 2 * jt.DIM_ROWS


def _both_refuse(unit, fact, code=abi.ERR_UNSUPPORTED, plan_only=True):
    with pytest.raises(oracle_lib.OracleError) as ei:
        (oracle_lib.plan if plan_only else oracle_lib.execute)(unit, fact)
    assert ei.value.code == code
    if plan_only:
        with pytest.raises(executor.QueryExecutionError) as ei2:
            executor.Executor().plan(unit, fact)
        assert ei2.value.code == code


def test_refused_joins(env):
    fact, dim, _ = env
    # range far wider than the row count: the reference switches to a baseline join table (PerfectJoinHashTable.cpp:235-246)
    sparse = abi.Table([(abi.kINT, True), (abi.kINT, False)])
    sparse.add_host_fragment([np.array([1, 1_000_000], dtype=np.int32), np.array([5, 6], dtype=np.int32)])
    _both_refuse(sqlmini.parse("SELECT COUNT(*) FROM t JOIN d ON t.fk32 = d.id;", fact, jt.FACT_NAMES, inner=(sparse, ["id", "a"])), fact)
    # duplicate inner keys: not one-to-one (found while filling the table)
    dup = abi.Table([(abi.kINT, True), (abi.kINT, False)])
    dup.add_host_fragment([np.array([1, 2, 2, 3], dtype=np.int32), np.array([5, 6, 7, 8], dtype=np.int32)])
    _both_refuse(sqlmini.parse("SELECT COUNT(*) FROM t JOIN d ON t.fk32 = d.id;", fact, jt.FACT_NAMES, inner=(dup, ["id", "a"])), fact, plan_only=False)
    # SEMI join (JoinType 2)
    b = abi.UnitBuilder(fact)
    b.join(dim, 0, 0)
    b.target(b.agg(abi.kCOUNT))
    b.unsupported["join_type"] = 2
    _both_refuse(b.build(), fact)
    # LEFT join whose inner ColumnVars were left NOT NULL: the planner would mis-size the key range
    b = abi.UnitBuilder(fact)
    b.join(dim, 0, 0)            # INNER while the columns are created ...
    b.group_by(0, 1)
    b.target(b.agg(abi.kCOUNT))
    b.unsupported["join_type"] = 1   # ... then declared LEFT
    _both_refuse(b.build(), fact, code=abi.ERR_INVALID_ARGUMENT)
    # a double as join key
    _both_refuse(sqlmini.parse("SELECT COUNT(*) FROM t JOIN d ON t.d = d.w;", fact, jt.FACT_NAMES, inner=(dim, jt.DIM_NAMES)), fact)


def test_malformed_inner_table_is_an_argument_error(env):
    """The inner table gets the checks the outer one gets: a descriptor with missing pieces is INVALID_ARGUMENT, never a
    dereference (the planner reads its chunk stats for the join table's range)."""
    import ctypes as C
    fact, dim, _ = env

    def broken(mutate):
        unit = parse("SELECT COUNT(*) FROM t JOIN d ON t.fk32 = d.id32;", fact, dim)
        info = unit._inner_built.info
        mutate(info)
        with pytest.raises(executor.QueryExecutionError) as ei:
            executor.Executor().plan(unit, fact)
        assert ei.value.code == abi.ERR_INVALID_ARGUMENT

    broken(lambda i: setattr(i.fragments[0], "col_stats", C.POINTER(abi.ChunkStats)()))
    broken(lambda i: setattr(i.fragments[0], "col_buffers", C.POINTER(C.c_void_p)()))
    broken(lambda i: setattr(i, "fragments", C.POINTER(abi.FragmentInfo)()))
    broken(lambda i: setattr(i, "col_types", C.POINTER(abi.TypeInfo)()))
    broken(lambda i: setattr(i.fragments[0], "num_tuples", -1))


def test_empty_tables():
    dim = jt.dim_table()
    empty_fact = jt.fact_table(0, seed=1, frag_rows=10)
    unit = parse("SELECT d.attr, COUNT(*) FROM t JOIN d ON t.fk32 = d.id32 GROUP BY d.attr;", empty_fact, dim)
    assert oracle_lib.execute(unit, empty_fact, entry_guess=16, has_card=True).rows() == []   # empty key range => baseline layout, like the single-table case
    fact = jt.fact_table(100, seed=1, frag_rows=40)
    empty_dim = abi.Table([(ty, nn) for _, ty, nn in jt.DIM_COLS])
    empty_dim.add_host_fragment([np.zeros(0, dtype=abi.NUMPY_OF[ty]) for _, ty, _ in jt.DIM_COLS])
    unit = parse("SELECT COUNT(*), SUM(t.v) FROM t JOIN d ON t.fk32 = d.id32;", fact, empty_dim)
    assert oracle_lib.execute(unit, fact).rows() == [(0, None)]
    assert executor.Executor().plan(unit, fact).as_dict() == oracle_lib.plan(unit, fact).as_dict()
