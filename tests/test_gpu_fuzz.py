"""Randomised query fuzzing on the GPU (`-m gpu`): random AND/OR filter trees, random target lists, random group key,
over the mixed-type / nullable random table of test_gpu_parity — CUDA path vs oracle, bit-exact integers."""
import random

import numpy as np
import pytest

import gpu_util as gu
from heavydb_b200 import abi, executor, sqlmini
from test_gpu_parity import RAND_COLS, RAND_NAMES, random_table

pytestmark = pytest.mark.gpu

INT_COLS = [n for n, t, _ in RAND_COLS if t not in (abi.kDOUBLE, abi.kFLOAT) and n != "sparse"]
FP_COLS = ["d", "dnn", "f32", "fnn"]      # DOUBLE and FLOAT (4-byte chunks, widened to double in the kernel)
FP_LIT = {"d": (-1500, 1500), "dnn": (-0.1, 1.1), "f32": (-120, 120), "fnn": (-1.5, 7.5)}
KEY_COLS = ["k8", "k16", "k32", "k64", "nn32", "nn64", "a8", "sparse"]
OPS = ["=", "<>", "<", ">", "<=", ">="]
LIT = {"k8": (-6, 21), "k16": (90, 410), "k32": (-1100, 1100), "k64": (10**9 - 10, 10**9 + 5010), "nn32": (-5, 305),
       "nn64": (-55, 55), "a8": (-130, 130), "a16": (-31000, 31000), "a32": (-2**31, 2**31), "a64": (-2**41, 2**41),
       "big": (-2**62, 2**62)}


def rand_cmp(rng):
    if rng.random() < 0.75:
        c = rng.choice(INT_COLS)
        lo, hi = LIT[c]
        lit = rng.randint(lo, hi)
        if rng.random() < 0.15:
            return f"{c} {rng.choice(OPS)} {lit + 0.5}"     # integer column against an fp literal
        return f"{c} {rng.choice(OPS)} {lit}"
    c = rng.choice(FP_COLS)
    lit = rng.uniform(*FP_LIT[c])
    return f"{c} {rng.choice(OPS)} {lit:.6f}"


def rand_leaf(rng):
    r = rng.random()
    if r > 0.9:          # column OP column
        a, b = rng.sample(INT_COLS + FP_COLS, 2)
        return f"{a} {rng.choice(OPS)} {b}"
    if r < 0.12:
        c = rng.choice(INT_COLS + FP_COLS)
        return f"{c} IS {'NOT ' if rng.random() < 0.5 else ''}NULL"
    if r < 0.2:
        c = rng.choice(INT_COLS)
        lo, hi = LIT[c]
        vals = ", ".join(str(rng.randint(lo, hi)) for _ in range(rng.randint(1, 3)))
        return f"{c} {'NOT ' if rng.random() < 0.3 else ''}IN ({vals})"
    return rand_cmp(rng)


def rand_cond(rng, depth=0):
    if depth >= 2 or rng.random() < 0.4:
        leaf = rand_leaf(rng)
        return f"NOT ({leaf})" if rng.random() < 0.1 else leaf
    if rng.random() < 0.1:
        return f"NOT {rand_cond_paren(rng, depth + 1)}"
    op = rng.choice(["AND", "OR"])
    a, b = rand_cond(rng, depth + 1), rand_cond(rng, depth + 1)
    return f"({a} {op} {b})"


def rand_cond_paren(rng, depth):
    return "(" + rand_cond(rng, depth) + ")"


def rand_query(rng, multi_key=False):
    key = rng.choice(KEY_COLS + [None, None])
    targets = [key] if key and rng.random() < 0.8 else []
    if multi_key:     # composite key: perfect hash over the cardinality product, or refused (baseline) when too large
        keys = rng.sample(["k8", "nn64", "a8", "k16", "nn32", "k32", "k64"], rng.choice([2, 2, 3]))
        key = ", ".join(keys)
        targets = [k for k in rng.sample(keys, len(keys)) if rng.random() < 0.8]
    for _ in range(rng.randint(1, 5)):
        agg = rng.choice(["COUNT", "SUM", "MIN", "MAX", "AVG", "COUNTSTAR"])
        if agg == "COUNTSTAR":
            targets.append("COUNT(*)")
        else:
            targets.append(f"{agg}({rng.choice(INT_COLS + FP_COLS)})")
    if all(not t.endswith(")") for t in targets):
        targets.append("COUNT(*)")
    sql = "SELECT " + ", ".join(targets) + " FROM r"
    n_conj = rng.choice([0, 1, 1, 2, 3])
    if n_conj:
        sql += " WHERE " + " AND ".join(rand_cond(rng) for _ in range(n_conj))
    if key:
        sql += f" GROUP BY {key}"
    return sql + ";"


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_multi_key_queries(seed):
    rng = random.Random(99 + seed)
    n = [3, 700, 30000, 90000][seed]
    table = random_table(n, seed=200 + seed, frag_rows=[2, 128, 30000, 25000][seed])
    dev = gu.DeviceTable(table)
    ran = 0
    for _ in range(60):
        sql = rand_query(rng, multi_key=True)
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            plan = executor.Executor().plan(unit, table)
        except executor.UnsupportedOnThisPath:
            continue     # cardinality product above the perfect-hash threshold, or an all-NULL key column
        force = abi.KERNEL_PERFECT_GLOBAL if rng.random() < 0.3 else 0
        try:
            gu.run_both(unit, table, dev_table=dev, force_kernel=force)
        except Exception as e:
            raise AssertionError(f"query: {sql}\nforce_kernel={force}\n{e}") from e
        ran += 1
    assert ran >= 15


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_queries(seed):
    rng = random.Random(1234 + seed)
    n = [1, 257, 5000, 40000, 40000, 120000][seed]
    table = random_table(n, seed=100 + seed, frag_rows=[10, 100, 1700, 40000, 9000, 50000][seed])
    dev = gu.DeviceTable(table)
    ran = 0
    for _ in range(60):
        sql = rand_query(rng)
        unit = sqlmini.parse(sql, table, RAND_NAMES)
        try:
            plan = executor.Executor().plan(unit, table, max_groups_buffer_entry_guess=3001, has_cardinality_estimation=True)
        except executor.UnsupportedOnThisPath:
            continue     # e.g. more slots than the ABI carries
        force = abi.KERNEL_PERFECT_GLOBAL if (plan.query_desc_type == abi.GroupByPerfectHash and rng.random() < 0.3) else 0
        try:
            gu.run_both(unit, table, entry_guess=3001, has_card=True, dev_table=dev, force_kernel=force)
        except Exception as e:
            raise AssertionError(f"query: {sql}\nforce_kernel={force}\n{e}") from e
        ran += 1
    assert ran >= 40


# ---- joins: random targets / filters / keys over both tables of the star schema -----------------------------
import join_tables as jt  # noqa: E402

J_INT = {"t.fk32": (-5, 1060), "t.x": (0, 100), "t.v": (-10**6, 10**6), "t.fk64": (-220, 1940),
         "d.attr": (0, 20), "d.attr8": (-100, 100), "d.big": (-2**50, 2**50), "d.id32": (3, 1003), "d.id64": (-100, 1900)}
J_FP = {"t.d": (0.0, 1.0), "d.w": (-30.0, 30.0)}
J_KEYS = ["d.attr", "d.attr8", "t.x", "d.id32", "d.big", None, None]
J_ON = ["t.fk32 = d.id32", "d.id32 = t.fk16", "t.fk64 = d.id64"]


def rand_join_leaf(rng):
    r = rng.random()
    if r > 0.9:          # column OP column, possibly across the two tables
        a, b = rng.sample(list(J_INT) + list(J_FP), 2)
        return f"{a} {rng.choice(OPS)} {b}"
    if r < 0.15:
        c = rng.choice(list(J_INT) + list(J_FP))
        return f"{c} IS {'NOT ' if rng.random() < 0.5 else ''}NULL"
    if r < 0.75:
        c = rng.choice(list(J_INT))
        lo, hi = J_INT[c]
        return f"{c} {rng.choice(OPS)} {rng.randint(lo, hi)}"
    c = rng.choice(list(J_FP))
    lo, hi = J_FP[c]
    return f"{c} {rng.choice(OPS)} {rng.uniform(lo, hi):.6f}"


def rand_join_query(rng):
    keys = []
    if rng.random() < 0.8:
        keys = [k for k in rng.sample(J_KEYS, rng.choice([1, 1, 2])) if k]
    targets = [k for k in keys if rng.random() < 0.8]
    for _ in range(rng.randint(1, 4)):
        agg = rng.choice(["COUNT", "SUM", "MIN", "MAX", "AVG", "COUNTSTAR"])
        targets.append("COUNT(*)" if agg == "COUNTSTAR" else f"{agg}({rng.choice(list(J_INT) + list(J_FP))})")
    if all(not t.endswith(")") for t in targets):
        targets.append("COUNT(*)")
    sql = "SELECT " + ", ".join(targets) + f" FROM t {'LEFT ' if rng.random() < 0.4 else ''}JOIN d ON {rng.choice(J_ON)}"
    n = rng.choice([0, 1, 1, 2])
    if n:
        conds = []
        for _ in range(n):
            a = rand_join_leaf(rng)
            if rng.random() < 0.4:
                a = f"({a} {rng.choice(['AND', 'OR'])} {rand_join_leaf(rng)})"
            if rng.random() < 0.15:
                a = f"NOT {a}" if a.startswith("(") else f"NOT ({a})"
            conds.append(a)
        sql += " WHERE " + " AND ".join(conds)
    if keys:
        sql += " GROUP BY " + ", ".join(keys)
    return sql + ";"


@pytest.mark.parametrize("seed", range(3))
def test_fuzz_join_queries(seed):
    rng = random.Random(4242 + seed)
    fact = jt.fact_table([50, 9000, 120000][seed], seed=60 + seed, frag_rows=[20, 2500, 50000][seed])
    dim = jt.dim_table(seed=7 + seed)
    dev = gu.DeviceTable(fact)
    ran = 0
    for _ in range(50):
        sql = rand_join_query(rng)
        unit = sqlmini.parse(sql, fact, jt.FACT_NAMES, inner=(dim, jt.DIM_NAMES))
        try:
            plan = executor.Executor().plan(unit, fact, max_groups_buffer_entry_guess=3001, has_cardinality_estimation=True)
        except executor.UnsupportedOnThisPath:
            continue
        force = abi.KERNEL_PERFECT_GLOBAL if (plan.query_desc_type == abi.GroupByPerfectHash and rng.random() < 0.3) else 0
        try:
            gu.run_both(unit, fact, entry_guess=3001, has_card=True, dev_table=dev, force_kernel=force)
        except Exception as e:
            raise AssertionError(f"query: {sql}\nforce_kernel={force}\n{e}") from e
        ran += 1
    assert ran >= 30
