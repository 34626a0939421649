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
