"""The planner's lowering of quals to the device filter program, checked on the CPU: tests/cpp/filter_emulator.cpp reads
the DevFilter of a planned query row by row with the semantics the scan kernel implements, and the number of passing
rows must be the oracle's COUNT(*) under the same WHERE clause (SQL three-valued logic, reference decoders)."""
import ctypes as C
import os
import random
import re
import subprocess

import pytest

import oracle_lib
import ref_tables as rt
import ref_time_table as tt
import sqlmini
import str_tables as stt
from heavydb_b200 import abi, build, executor
from test_gpu_fuzz import rand_query
from test_gpu_parity import RAND_NAMES, random_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    build.build()
    so = tmp_path_factory.mktemp("emu") / "libfilter_emulator.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "filter_emulator.cpp"), "-o", str(so)])
    E = C.CDLL(str(so))
    E.b2q_test_eval_filter.restype = C.c_int32
    E.b2q_test_eval_filter.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64]
    return E


def passing_rows(emu, unit, table):
    """Plan on the host, then run the lowered filter over every row of every fragment."""
    L = executor.lib()
    bt = table.build(abi.CPU_LEVEL)
    co, eo = executor.compilation_options(), executor.execution_options()
    h = C.c_void_p()
    rc = L.b2q_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(co), C.byref(eo), 0, 0, C.byref(h))
    assert rc == 0, L.b2q_last_error_message()
    n = 0
    try:
        for f in table.fragments:
            ptrs = (C.c_void_p * table.num_cols)(*[a.ctypes.data if a is not None else None for a in f.host_cols])
            for row in range(f.num_tuples):
                r = emu.b2q_test_eval_filter(h, ptrs, row)
                assert r >= 0
                n += r
    finally:
        L.b2q_query_free(h)
    return n


def where_of(sql):
    m = re.search(r" WHERE (.*?)( GROUP BY | ORDER BY |;)", sql)
    return m.group(1) if m else None


def check(emu, table, names, tname, where):
    sql = f"SELECT COUNT(*) FROM {tname} WHERE {where};"
    unit = sqlmini.parse(sql, table, names)
    want = oracle_lib.execute(unit, table).rows()[0][0]
    assert passing_rows(emu, unit, table) == want, sql


@pytest.mark.parametrize("seed", range(3))
def test_random_filter_trees(emu, seed):
    rng = random.Random(4400 + seed)
    table = random_table([60, 700, 700][seed], seed=50 + seed, frag_rows=[13, 200, 700][seed])
    checked = 0
    for _ in range(150):
        where = where_of(rand_query(rng))
        if not where:
            continue
        try:
            check(emu, table, RAND_NAMES, "r", where)
        except executor.UnsupportedOnThisPath:
            continue
        checked += 1
    assert checked >= 80


def test_dictionary_time_and_days_encoded_columns(emu):
    table = stt.str_table(1500, seed=8, frag_rows=400)
    wheres = {where_of(q) for q in stt.STR_QUERIES if where_of(q)}
    wheres |= {"dd = 1555286400", "dd <> 1555286400", "dd = 1555286401", "dd <> 1555286401", "dd < 1555286401", "dd <= 1555286399",
               "dd > 1555286399 AND NOT (dd >= 1555459200)", "dd16 = 864000000 OR dd16 < 863049600 OR dd IS NULL",
               "NOT (dd16 <= 864000001) AND dd IS NOT NULL", "dd16 > dd OR dd16 = dd16", "ts <= 1600000100 AND s8 <> 200 AND NOT (s16 = 7)",
               "dd BETWEEN 1555200000 AND 1556000000", "dd NOT IN (1555286400, 1555372800, 5)", "dd16 >= -9000000000000 AND dd16 <= 9000000000000"}
    for w in sorted(wheres):
        check(emu, table, stt.STR_NAMES, "s", w)


def test_golden_time_table(emu):
    table = tt.make_table(tt.time_rows())
    for q in tt.TIME_QUERIES:
        if where_of(q):
            check(emu, table, tt.TIME_NAMES, "test", where_of(q))


def n_terms(emu, unit, table):
    L = executor.lib()
    bt = table.build(abi.CPU_LEVEL)
    co, eo = executor.compilation_options(), executor.execution_options()
    h = C.c_void_p()
    assert L.b2q_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(co), C.byref(eo), 0, 0, C.byref(h)) == 0
    emu.b2q_test_filter_terms.restype = C.c_int32
    emu.b2q_test_filter_terms.argtypes = [C.c_void_p]
    n = emu.b2q_test_filter_terms(h)
    L.b2q_query_free(h)
    return n


def test_in_lists_fold_consecutive_values_into_ranges(emu):
    """`c IN (...)` is an OR chain of equalities (NOT IN: an AND chain of <>): runs of consecutive values become ONE range
    term, so a dense 30-value list — past the 16-leaf program as written — is two terms here."""
    table = random_table(900, seed=61, frag_rows=250)
    dense = ", ".join(str(v) for v in range(-3, 27))
    cases = [
        (f"k8 IN ({dense})", 1), (f"k8 NOT IN ({dense})", 1), ("k8 IN (1, 2, 3, 7, 8, 20)", 3), ("NOT (k8 IN (5, 4, 3) OR k16 = 7)", 2),
        ("k16 IN (100, 101, 102) AND a8 NOT IN (-1, 0, 1, 2)", 2), ("nn32 IN (3, 4, 5) OR nn32 IN (6, 7) OR d < 0.5", 2),
        ("k8 = 2 OR (k32 < 10 AND k16 > 5) OR k8 = 3 OR k8 = 4 OR a16 IS NULL", 4), ("k64 IN (1000000001, 1000000002, 1000000004)", 2),
        ("a8 IN (5, 5, 6) OR a8 = 127", 2), ("NOT (a16 <> 10 AND a16 <> 11 AND a16 <> 12)", 1),
    ]
    for where, terms in cases:
        check(emu, table, RAND_NAMES, "r", where)
        assert n_terms(emu, sqlmini.parse(f"SELECT COUNT(*) FROM r WHERE {where};", table, RAND_NAMES), table) == terms, where
    s = stt.str_table(1500, seed=9, frag_rows=400)
    for where, terms in [("dd IN (1555286400, 1555372800, 1555459200)", 1), ("dd16 NOT IN (864000000, 864086400, 863913600, 5)", 2),
                         ("s8 IN (3, 4, 5, 6, 200) OR str = 7 OR str = 8", 3), ("dt IN (1555200000, 1555286400) OR dt = 1555372801", 3)]:
        check(emu, s, stt.STR_NAMES, "s", where)
        assert n_terms(emu, sqlmini.parse(f"SELECT COUNT(*) FROM s WHERE {where};", s, stt.STR_NAMES), s) == terms, where


@pytest.mark.parametrize("seed", range(2))
def test_random_dense_in_lists(emu, seed):
    rng = random.Random(880 + seed)
    table = random_table(600, seed=70 + seed, frag_rows=170)
    cols = {"k8": (-2, 12), "k16": (90, 140), "nn32": (0, 40), "a8": (-128, 127), "k64": (1000000000, 1000000040), "nn64": (-50, 50)}
    for _ in range(120):
        parts = []
        for _ in range(rng.randint(1, 3)):
            c = rng.choice(sorted(cols))
            lo, hi = cols[c]
            start = rng.randint(lo, hi)
            vals = [start + i for i in range(rng.randint(1, 6))] + [rng.randint(lo, hi) for _ in range(rng.randint(0, 2))]
            rng.shuffle(vals)
            parts.append(f"{c} {'NOT IN' if rng.random() < 0.4 else 'IN'} ({', '.join(map(str, vals))})")
        where = parts[0]
        for p in parts[1:]:
            where = f"({where}) {rng.choice(['AND', 'OR'])} {'NOT ' if rng.random() < 0.2 else ''}({p})"
        try:
            check(emu, table, RAND_NAMES, "r", where)
        except executor.UnsupportedOnThisPath:
            continue


def test_range_leaves_on_one_column_fold_into_one_term(emu):
    """BETWEEN (two conjuncts, possibly two separate quals), redundant bounds, `=` with bounds, NOT BETWEEN (an OR of two
    one-sided leaves): one range test per column."""
    table = random_table(900, seed=62, frag_rows=250)
    cases = [
        ("k16 BETWEEN 100 AND 130", 1), ("k16 >= 100 AND k16 <= 130 AND k16 < 125 AND k16 > 90", 1),
        ("k16 NOT BETWEEN 100 AND 130", 1), ("NOT (k16 >= 100 AND k16 <= 130)", 1), ("k16 < 100 OR k16 > 130 OR k16 = 110", 2),
        ("a8 > 5 AND d < 0.5 AND a8 <= 60 AND nn32 <> 7", 3), ("k8 = 3 AND k8 < 10 AND k8 > 0", 1), ("k8 = 3 AND k8 > 5", 1),
        ("k8 >= 2 AND k8 <= 9 AND k8 NOT IN (4, 5, 6)", 2), ("(a16 > 100 AND a16 < 20000) OR (a16 > -20000 AND a16 < -100)", 2),
        ("k64 > 1000000010 AND k64 <= 1000000400 AND k32 BETWEEN -50 AND 50 AND a64 IS NOT NULL", 3),
        ("a16 < 100 OR a16 > 99", 1), ("a16 <= 5 OR a16 >= 6", 1), ("NOT (a32 BETWEEN -1000000 AND 1000000) AND a32 < 2000000000", 2),
    ]
    for where, terms in cases:
        check(emu, table, RAND_NAMES, "r", where)
        assert n_terms(emu, sqlmini.parse(f"SELECT COUNT(*) FROM r WHERE {where};", table, RAND_NAMES), table) == terms, where
    s = stt.str_table(1500, seed=10, frag_rows=400)
    for where, terms in [("dd BETWEEN 1555200000 AND 1556000000", 1), ("dd > 1555286399 AND dd < 1555459201 AND dd16 >= 863308800", 2),
                         ("dd16 NOT BETWEEN 863913601 AND 864086399", 1), ("ts >= 1600000050 AND ts < 1600000200 AND dt <= 1556000000 AND dt > 1555200001", 2)]:
        check(emu, s, stt.STR_NAMES, "s", where)
        assert n_terms(emu, sqlmini.parse(f"SELECT COUNT(*) FROM s WHERE {where};", s, stt.STR_NAMES), s) == terms, where


@pytest.mark.parametrize("seed", range(2))
def test_random_range_chains(emu, seed):
    rng = random.Random(990 + seed)
    table = random_table(600, seed=75 + seed, frag_rows=170)
    cols = {"k8": (-2, 12), "k16": (90, 140), "nn32": (0, 40), "a8": (-128, 127), "k64": (1000000000, 1000000040), "a16": (-32768, 32767), "big": (-2**62, 2**62)}
    for _ in range(150):
        leaves = []
        for _ in range(rng.randint(2, 6)):
            c = rng.choice(sorted(cols))
            lo, hi = cols[c]
            leaves.append(f"{'NOT ' if rng.random() < 0.15 else ''}({c} {rng.choice(['<', '<=', '>', '>=', '=', '<>'])} {rng.randint(lo, hi)})")
        where = leaves[0]
        for lf in leaves[1:]:
            where = f"{where} {rng.choice(['AND', 'AND', 'OR'])} {lf}" if rng.random() < 0.7 else f"({where}) {rng.choice(['AND', 'OR'])} {lf}"
        if rng.random() < 0.2:
            where = f"NOT ({where})"
        try:
            check(emu, table, RAND_NAMES, "r", where)
        except executor.UnsupportedOnThisPath:
            continue


def test_double_bounds_fold_too(emu):
    """Bounding boxes: `d BETWEEN a AND b` on a DOUBLE column is one closed-range test; `<` / `>` step to the neighbouring
    double; `d < a OR d > b` is the negated range; NULL_DOUBLE never passes."""
    table = random_table(900, seed=63, frag_rows=250)
    cases = [
        ("d BETWEEN -500.5 AND 700.25", 1), ("d > -500.5 AND d < 700.25 AND dnn >= 0.25 AND dnn <= 0.75", 2),
        ("d NOT BETWEEN -500.5 AND 700.25", 1), ("d < -500 OR d > 700", 1), ("dnn > 0.5 AND dnn < 0.5", 1), ("dnn >= 0.5 AND dnn <= 0.5", 1),
        ("d >= 0 AND d < 1000 AND k8 BETWEEN 2 AND 9 AND d > 10", 2), ("NOT (d >= -100 AND d <= 100) AND d IS NOT NULL", 2),
        ("d < 5 OR d >= 5", 1), ("d <= 5 OR d >= 5", 2), ("d > 100 AND d > 200 AND d > 5", 1), ("dnn < 0.3 OR dnn > 0.6 OR dnn = 0.45", 2),
    ]
    for where, terms in cases:
        check(emu, table, RAND_NAMES, "r", where)
        assert n_terms(emu, sqlmini.parse(f"SELECT COUNT(*) FROM r WHERE {where};", table, RAND_NAMES), table) == terms, where
    rng = random.Random(5)
    for _ in range(150):
        leaves = [f"{'NOT ' if rng.random() < 0.15 else ''}({rng.choice(['d', 'dnn', 'd'])} {rng.choice(['<', '<=', '>', '>=', '=', '<>'])} {rng.choice([round(rng.uniform(-1500, 1500), 3), round(rng.random(), 4), rng.randint(-5, 5)])})"
                  for _ in range(rng.randint(2, 5))]
        where = leaves[0]
        for lf in leaves[1:]:
            where = f"{where} {rng.choice(['AND', 'AND', 'OR'])} {lf}" if rng.random() < 0.7 else f"({where}) {rng.choice(['AND', 'OR'])} {lf}"
        check(emu, table, RAND_NAMES, "r", where)


def test_is_not_null_next_to_a_comparison_is_dropped(emu):
    """Calcite adds `c IS NOT NULL` next to filters and join keys; a comparison on c in the same AND chain already fails on
    NULL, so the leaf costs nothing on the device."""
    table = random_table(900, seed=64, frag_rows=250)
    cases = [("a16 IS NOT NULL AND a16 > 100", 1), ("a16 > 100 AND k16 < 120 AND NOT (a16 IS NULL)", 2), ("d IS NOT NULL AND d < 0.5 AND a8 IS NOT NULL", 2),
             ("a16 IS NOT NULL AND a16 <> 5", 1), ("a16 IS NOT NULL OR a16 > 100", 2), ("a16 IS NULL AND a16 > 100", 2),
             ("(a16 IS NOT NULL AND k8 = 3) OR (a8 IS NOT NULL AND a8 BETWEEN 1 AND 50)", 3), ("a64 IS NOT NULL AND a64 IN (5, 6, 7) AND nn64 IS NOT NULL", 3)]
    for where, terms in cases:
        check(emu, table, RAND_NAMES, "r", where)
        assert n_terms(emu, sqlmini.parse(f"SELECT COUNT(*) FROM r WHERE {where};", table, RAND_NAMES), table) == terms, where


def test_deleted_rows_term_precedes_the_folded_chain(emu):
    """A $deleted$ column adds one term in front of the quals (codegenSkipDeletedOuterTableRow); the folded chain starts one
    stack slot higher and is AND-ed to it."""
    import numpy as np
    rng = np.random.default_rng(12)
    n = 800
    t = abi.Table([(abi.kINT, False), (abi.kDOUBLE, False), (abi.kBOOLEAN, True), (abi.kBIGINT, True)], deleted_column=2)
    for b in range(0, n, 300):
        m = min(300, n - b)
        x = rng.integers(-50, 50, m).astype(np.int32)
        x[rng.random(m) < 0.1] = abi.NULL_INT
        d = rng.normal(0, 10, m)
        d[rng.random(m) < 0.1] = abi.NULL_DOUBLE
        t.add_host_fragment([x, d, (rng.random(m) < 0.3).astype(np.int8), rng.integers(0, 1000, m).astype(np.int64)])
    names = ["x", "d", "del", "v"]
    for where in ["x BETWEEN -10 AND 10", "x >= -10 AND x <= 10 AND d > -5 AND d < 5", "x IN (1, 2, 3) OR (d < -3 AND v BETWEEN 100 AND 900 AND v <> 500)",
                  "x IS NOT NULL AND x > 0 AND NOT (d BETWEEN -1 AND 1)", "(x < -20 OR x > 20) AND (v IN (5, 6, 7, 8) OR d >= 0)", "x = 5"]:
        check(emu, t, names, "t", where)


def test_join_queries_read_the_same_filter_through_the_join_index(emu):
    """INNER / LEFT star joins: the probe is done here in numpy (inner columns gathered at the matching row, the chunk's NULL
    where a LEFT join has no match), the lowered filter — with the LEFT join's nullable inner columns — runs on that
    denormalised row, and the count must be the oracle's."""
    import numpy as np
    import join_tables as jt
    emu.b2q_test_eval_filter_joined.restype = C.c_int32
    emu.b2q_test_eval_filter_joined.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64]
    fact, dim = jt.fact_table(1200, seed=31, frag_rows=500), jt.dim_table(seed=11)
    dcols = dim.fragments[0].host_cols
    ids = dcols[jt.DIM_NAMES.index("id32")]
    pos = {int(k): i for i, k in enumerate(ids)}
    L = executor.lib()
    wheres = ["t.x BETWEEN 10 AND 60 AND d.attr8 >= -50 AND d.attr8 < 50", "d.attr IN (1, 2, 3, 9) OR t.x < 5", "d.w > -5 AND d.w <= 5 AND d.attr IS NOT NULL AND d.attr <> 4",
              "NOT (d.attr8 BETWEEN -20 AND 20) AND t.v IS NOT NULL AND t.v > 0", "d.big > 0 AND d.big < 900000000000000 AND t.x <> 7 AND t.x <> 8 AND t.x <> 9",
              "d.attr8 < t.x OR d.attr8 IN (5, 6, 7)", "d.attr IS NULL OR d.attr >= 10"]
    for left in (False, True):
        for where in wheres:
            sql = f"SELECT COUNT(*) FROM t {'LEFT ' if left else ''}JOIN d ON t.fk32 = d.id32 WHERE {where};"
            unit = sqlmini.parse(sql, fact, jt.FACT_NAMES, inner=(dim, jt.DIM_NAMES))
            want = oracle_lib.execute(unit, fact).rows()[0][0]
            bt = fact.build(abi.CPU_LEVEL)
            co, eo = executor.compilation_options(), executor.execution_options()
            h = C.c_void_p()
            assert L.b2q_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(co), C.byref(eo), 0, 0, C.byref(h)) == 0, L.b2q_last_error_message()
            got = 0
            for f in fact.fragments:
                fk = f.host_cols[jt.FACT_NAMES.index("fk32")]
                idx = np.array([pos.get(int(k), -1) if k != abi.NULL_INT else -1 for k in fk], dtype=np.int64)
                gathered = []
                for c, a in enumerate(dcols):
                    g = a[np.maximum(idx, 0)].copy()
                    g[idx < 0] = dim.physical_null(c)
                    gathered.append(g)
                arrays = list(f.host_cols) + gathered
                ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
                for row in range(f.num_tuples):
                    if not left and idx[row] < 0:
                        continue          # INNER: no match, no row
                    r = emu.b2q_test_eval_filter_joined(h, ptrs, row)
                    assert r >= 0
                    got += r
            L.b2q_query_free(h)
            assert got == want, sql


def per_entry_counts(emu, unit, table):
    """COUNT(*) per ENTRY from the lowered program: filter, then the key mapping (perfect hash)."""
    import numpy as np
    emu.b2q_test_group_index.restype = C.c_int64
    emu.b2q_test_group_index.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64]
    L = executor.lib()
    bt = table.build(abi.CPU_LEVEL)
    co, eo = executor.compilation_options(), executor.execution_options()
    h = C.c_void_p()
    assert L.b2q_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(co), C.byref(eo), 0, 0, C.byref(h)) == 0, L.b2q_last_error_message()
    plan = L.b2q_query_plan(h).contents
    counts = np.zeros(plan.entry_count, dtype=np.int64)
    for f in table.fragments:
        ptrs = (C.c_void_p * table.num_cols)(*[a.ctypes.data if a is not None else None for a in f.host_cols])
        for row in range(f.num_tuples):
            if emu.b2q_test_eval_filter(h, ptrs, row) == 1:
                e = emu.b2q_test_group_index(h, ptrs, row)
                assert e >= 0, (e, row)
                counts[e] += 1
    L.b2q_query_free(h)
    return counts


def oracle_count_column(res, slot):
    """The COUNT slot of every entry of the oracle's (row-wise) buffer."""
    import numpy as np
    p = res.plan
    buf = res.buffer()
    w = p.slot_padded_width[slot]
    rows = buf.reshape(p.entry_count, p.row_size)
    col = rows[:, p.slot_offset[slot]:p.slot_offset[slot] + w].copy()
    return col.view(np.int32 if w == 4 else np.int64).reshape(-1).astype(np.int64)


def test_key_mapping_places_every_row_in_the_oracles_entry(emu):
    """GROUP BY keys -> entry index as the planner lowers it (min, NULL entry, mixed radix, the DATE day bucket with a min off
    the day grid, days-encoded chunks indexed in days): per-entry COUNT(*) equals the oracle's buffer, entry by entry."""
    cases = [
        (random_table(900, seed=65, frag_rows=250), RAND_NAMES, "r",
         ["SELECT k8, COUNT(*) FROM r GROUP BY k8;", "SELECT k16, COUNT(*) FROM r WHERE a8 > 0 GROUP BY k16;", "SELECT nn64, COUNT(*) FROM r WHERE nn64 > -20 GROUP BY nn64;",
          "SELECT k8, k16, COUNT(*) FROM r WHERE d < 500 GROUP BY k8, k16;", "SELECT nn32, k8, COUNT(*) FROM r GROUP BY nn32, k8;", "SELECT k64, COUNT(*) FROM r WHERE k64 >= 1000000100 GROUP BY k64;"]),
        (stt.str_table(1500, seed=11, frag_rows=400), stt.STR_NAMES, "s",
         ["SELECT dd, COUNT(*) FROM s GROUP BY dd;", "SELECT dd16, COUNT(*) FROM s WHERE dd > 1555286400 GROUP BY dd16;", "SELECT dt, COUNT(*) FROM s WHERE dt > 1555286410 GROUP BY dt;",
          "SELECT dd, dt, COUNT(*) FROM s WHERE dd <= 1557000000 GROUP BY dd, dt;", "SELECT s8, COUNT(*) FROM s GROUP BY s8;", "SELECT str, s8, COUNT(*) FROM s GROUP BY str, s8;",
          "SELECT ts, COUNT(*) FROM s GROUP BY ts;"]),
        (tt.make_table(tt.time_rows()), tt.TIME_NAMES, "test",
         ["SELECT o, COUNT(*) FROM test WHERE o <= 936835200 GROUP BY o;", "SELECT fx, COUNT(*) FROM test GROUP BY fx;", "SELECT o1, o2, COUNT(*) FROM test GROUP BY o1, o2;", "SELECT m, COUNT(*) FROM test GROUP BY m;"]),
    ]
    for table, names, _tname, sqls in cases:
        for sql in sqls:
            unit = sqlmini.parse(sql, table, names)
            res = oracle_lib.execute(unit, table)
            assert res.plan.query_desc_type == abi.GroupByPerfectHash and not res.plan.output_columnar, sql
            n_keys = max(res.plan.num_group_cols, 1)
            got = per_entry_counts(emu, unit, table)
            want = oracle_count_column(res, slot=n_keys)       # targets: the key column(s), then COUNT(*)
            assert (got == want).all(), sql


def run_program(emu, unit, table, output_columnar=False, entry_guess=0, has_card=False):
    """The whole lowered program on the host (tests/cpp/filter_emulator.cpp: b2q_test_run_program) -> result buffer bytes."""
    import numpy as np
    emu.b2q_test_run_program.restype = C.c_int32
    emu.b2q_test_run_program.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.c_int64), C.c_void_p]
    L = executor.lib()
    bt = table.build(abi.CPU_LEVEL)
    co, eo = executor.compilation_options(), executor.execution_options(output_columnar_hint=output_columnar)
    h = C.c_void_p()
    rc = L.b2q_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(co), C.byref(eo), entry_guess, int(has_card), C.byref(h))
    assert rc == 0, L.b2q_last_error_message()
    plan = L.b2q_query_plan(h).contents
    nf = len(table.fragments)
    keep = [(C.c_void_p * table.num_cols)(*[a.ctypes.data if a is not None else None for a in f.host_cols]) for f in table.fragments]
    frag_cols = (C.POINTER(C.c_void_p) * max(nf, 1))(*[C.cast(k, C.POINTER(C.c_void_p)) for k in keep])
    frag_rows = (C.c_int64 * max(nf, 1))(*[f.num_tuples for f in table.fragments])
    out = np.zeros(max(plan.buffer_size, 8), dtype=np.uint8)
    rc = emu.b2q_test_run_program(h, nf, frag_cols, frag_rows, out.ctypes.data)
    L.b2q_query_free(h)
    return rc, out[:plan.buffer_size]


def assert_buffers_match(got, want, sql, float_atol=0.05):
    """Bit-exact, except that 8-byte words which read as doubles may differ in the last bits (fp summation order)."""
    import numpy as np
    if np.array_equal(got, want):
        return
    assert got.size == want.size and got.size % 8 == 0, sql
    g, w = got.view(np.int64), want.view(np.int64)
    bad = np.nonzero(g != w)[0]
    gd, wd = got.view(np.float64)[bad], want.view(np.float64)[bad]
    with np.errstate(invalid="ignore"):
        as_double = np.isfinite(gd) & (np.abs(gd - wd) <= 1e-9 * np.abs(wd))
    # a SUM(FLOAT) slot: float bits in the low half (the oracle adds in float like agg_sum_float, the program adds the
    # widened values in double and narrows once), the init value's high half untouched
    gf, wf = got.view(np.float32)[0::2][bad].astype(np.float64), want.view(np.float32)[0::2][bad].astype(np.float64)
    same_hi = got.view(np.int32)[1::2][bad] == want.view(np.int32)[1::2][bad]
    as_float = same_hi & np.isfinite(gf) & (np.abs(gf - wf) <= rt.FLOAT_SUM_RTOL * np.abs(wf) + float_atol)
    assert np.all(as_double | as_float), (sql, bad[:5], g[bad][:5], w[bad][:5])


def test_whole_program_reproduces_the_oracles_buffer(emu):
    """filter -> entry -> accumulators with their NULL-skip rules -> materialise, all read from the lowered program on the
    host: the result buffer is the oracle's (row-wise and columnar; keyless and keyed; single, composite and DATE keys;
    non-grouped)."""
    from test_gpu_parity import RAND_QUERIES
    import reduce_ladder as rl
    ran = 0
    cases = [(random_table(1200, seed=66, frag_rows=350), RAND_NAMES, list(RAND_QUERIES)),
             (stt.str_table(1500, seed=12, frag_rows=400), stt.STR_NAMES, [q for q in stt.STR_QUERIES if " ORDER BY " not in q]),
             (tt.make_table(tt.time_rows()), tt.TIME_NAMES, [q for q in tt.TIME_QUERIES if " ORDER BY " not in q]),
             (rl.ladder_table(241, 7, 17, 60, seed=5)[0], rl.NAMES, [rl.QUERY])]
    for table, names, sqls in cases:
        for sql in sqls:
            unit = sqlmini.parse(sql, table, names)
            for columnar in (False, True):
                try:
                    res = oracle_lib.execute(unit, table, entry_guess=3001, has_card=True, output_columnar=columnar)
                except oracle_lib.OracleError:
                    continue
                if res.plan.query_desc_type not in (abi.GroupByPerfectHash, abi.NonGroupedAggregate):
                    continue
                rc, got = run_program(emu, unit, table, output_columnar=columnar, entry_guess=3001, has_card=True)
                assert rc == 0, (sql, rc)
                assert_buffers_match(got, res.buffer(), sql)
                ran += 1
    assert ran >= 60


@pytest.mark.parametrize("seed", range(3))
def test_random_queries_whole_program(emu, seed):
    """The GPU fuzz generator's queries (filters, single / composite keys, every aggregate over every column type), planned
    on the host and run through the host reading of the lowered program: same buffer as the oracle, both layouts."""
    rng = random.Random(31000 + seed)
    table = random_table([40, 900, 2000][seed], seed=520 + seed, frag_rows=[11, 300, 2000][seed])
    ran = 0
    for i in range(70):
        sql = rand_query(rng, multi_key=(i % 3 == 0))
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        for columnar in (False, True):
            try:
                res = oracle_lib.execute(unit, table, entry_guess=6000, has_card=True, output_columnar=columnar)
            except oracle_lib.OracleError:
                continue
            if res.plan.query_desc_type not in (abi.GroupByPerfectHash, abi.NonGroupedAggregate):
                continue
            rc, got = run_program(emu, unit, table, output_columnar=columnar, entry_guess=6000, has_card=True)
            assert rc == 0, sql
            assert_buffers_match(got, res.buffer(), sql)
            ran += 1
    assert ran >= 40


def run_join_program(emu, unit, fact, dim, left, entry_guess=4000):
    """Probe in numpy, then the whole lowered join program on the denormalised rows -> (rc, result buffer bytes)."""
    import numpy as np
    emu.b2q_test_run_program_joined.restype = C.c_int32
    emu.b2q_test_run_program_joined.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.c_int64),
                                                C.POINTER(C.POINTER(C.c_uint8)), C.c_void_p]
    L = executor.lib()
    dcols = dim.fragments[0].host_cols
    e = unit.unit.exprs[unit.unit.join_qual]
    a, b = unit.unit.exprs[e.left], unit.unit.exprs[e.right]
    outer, inner = (a, b) if a.rte_idx == 0 else (b, a)
    ikeys = dcols[inner.col_id]
    inull = dim.physical_null(inner.col_id)
    pos = {int(k): i for i, k in enumerate(ikeys) if dim.col_types[inner.col_id][1] or k != inull}
    bt = fact.build(abi.CPU_LEVEL)
    co, eo = executor.compilation_options(), executor.execution_options()
    h = C.c_void_p()
    assert L.b2q_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(co), C.byref(eo), entry_guess, 1, C.byref(h)) == 0, L.b2q_last_error_message()
    plan = L.b2q_query_plan(h).contents
    keep, valids = [], []
    for f in fact.fragments:
        fk = f.host_cols[outer.col_id]
        onull = fact.physical_null(outer.col_id)
        idx = np.array([-1 if (not fact.col_types[outer.col_id][1] and k == onull) else pos.get(int(k), -1) for k in fk], dtype=np.int64)
        gathered = []
        for c, col in enumerate(dcols):
            g = col[np.maximum(idx, 0)].copy() if len(col) else np.full(len(fk), dim.physical_null(c), dtype=col.dtype)
            g[idx < 0] = dim.physical_null(c)
            gathered.append(g)
        arrays = list(f.host_cols) + gathered
        keep.append((arrays, (C.c_void_p * len(arrays))(*[x.ctypes.data for x in arrays])))
        valids.append(np.ascontiguousarray(np.ones(len(fk), dtype=np.uint8) if left else (idx >= 0).astype(np.uint8)))
    nf = len(keep)
    frag_cols = (C.POINTER(C.c_void_p) * nf)(*[C.cast(k[1], C.POINTER(C.c_void_p)) for k in keep])
    frag_rows = (C.c_int64 * nf)(*[f.num_tuples for f in fact.fragments])
    frag_valid = (C.POINTER(C.c_uint8) * nf)(*[v.ctypes.data_as(C.POINTER(C.c_uint8)) for v in valids])
    out = np.zeros(max(plan.buffer_size, 8), dtype=np.uint8)
    rc = emu.b2q_test_run_program_joined(h, nf, frag_cols, frag_rows, frag_valid, out.ctypes.data)
    L.b2q_query_free(h)
    return rc, out[:plan.buffer_size]


def test_join_programs_reproduce_the_oracles_buffer(emu):
    """INNER and LEFT star joins: probe in numpy, then the whole lowered program on the denormalised rows — inner columns as
    keys, filter operands and aggregate arguments, nullable under a LEFT join — gives the oracle's buffer."""
    import join_tables as jt
    from test_gpu_fuzz import rand_join_query
    fact, dim = jt.fact_table(1500, seed=41, frag_rows=600), jt.dim_table(seed=13)
    rng = random.Random(77)
    ran = 0
    for sql in jt.JOIN_QUERIES + jt.LEFT_JOIN_QUERIES + [rand_join_query(rng) for _ in range(60)]:
        unit = sqlmini.parse(sql, fact, jt.FACT_NAMES, inner=(dim, jt.DIM_NAMES))
        if unit.unit.num_order_entries:
            continue
        try:
            res = oracle_lib.execute(unit, fact, entry_guess=4000, has_card=True)
        except oracle_lib.OracleError:
            continue
        if res.plan.query_desc_type not in (abi.GroupByPerfectHash, abi.NonGroupedAggregate):
            continue
        rc, got = run_join_program(emu, unit, fact, dim, left=" LEFT JOIN " in sql)
        assert rc == 0, (sql, rc)
        assert_buffers_match(got, res.buffer(), sql)
        ran += 1
    assert ran >= 30
