"""The reference's golden table `test` (Tests/ExecuteTest.cpp:142-186 schema, :30063-30115 INSERT templates) with its
dictionary-encoded string columns and its fixed-width integer column next to the numeric columns of ref_tables.py — the columns the
path accepts: str (varchar(10), dictionary by default), null_str, fixed_str / fixed_null_str (DICT(16): uint16 ids, NULL = 65535),
shared_dict, ss (TEXT ENCODING DICT), fx (INT ENCODING FIXED(16)).  Dictionary ids are assigned in insertion order
(StringDictionary::getOrAdd), so 'foo' = 0, 'bar' = 1, 'baz' = 2 in str; results carry ids and are translated back for the comparison
with SQLite, which holds the strings."""
import sqlite3

import numpy as np

import ref_tables as rt
from heavydb_b200 import abi

# name, type, notnull, physical bytes (0 = the type's own)
EXTRA_COLS = [
    ("str", abi.kVARCHAR, False, 0),
    ("null_str", abi.kTEXT, False, 0),
    ("fixed_str", abi.kTEXT, False, 2),
    ("fixed_null_str", abi.kTEXT, False, 2),
    ("shared_dict", abi.kTEXT, False, 0),
    ("ss", abi.kTEXT, False, 0),
    ("fx", abi.kINT, False, 2),
]
FULL_COLS = [(n, t, nn, 0) for n, t, nn in rt.TEST_COLS] + EXTRA_COLS
FULL_NAMES = [c[0] for c in FULL_COLS]
# per INSERT template (x10, x5, x5), in EXTRA_COLS order
_E1 = ("foo", None, "foo", None, "foo", "fish", 9)
_E2 = ("bar", None, "bar", None, None, None, None)
_E3 = ("baz", None, None, None, "baz", "boat", 11)
_STR_DICT = ["foo", "bar", "baz"]
DICTS = {"str": _STR_DICT, "null_str": [], "fixed_str": ["foo", "bar"], "fixed_null_str": [],
         "shared_dict": _STR_DICT,                  # SHARED DICTIONARY (shared_dict) REFERENCES test(str) (ExecuteTest.cpp:30049): ONE dictionary
         "ss": ["fish", "boat"]}


def full_rows(num_rows: int = rt.G_NUM_ROWS):
    return [rt._T1 + _E1] * num_rows + [rt._T2 + _E2] * (num_rows // 2) + [rt._T3 + _E3] * (num_rows // 2)


def make_table(rows, fragment_size: int = 2) -> abi.Table:
    t = abi.Table([(ty, nn) for _, ty, nn, _ in FULL_COLS], encoded_sizes=[e for *_, e in FULL_COLS])
    arrays = []
    for c, (name, ty, _nn, _e) in enumerate(FULL_COLS):
        null = t.physical_null(c)
        if name in DICTS:
            vals = [null if r[c] is None else DICTS[name].index(r[c]) for r in rows]
        else:
            vals = [null if r[c] is None else r[c] for r in rows]
        arrays.append(np.array(vals, dtype=t.physical_dtype(c)))
    for b in range(0, len(rows), fragment_size):
        t.add_host_fragment([a[b:b + fragment_size] for a in arrays])
    return t


def make_sqlite(rows, name="test"):
    con = sqlite3.connect(":memory:")
    decl = ", ".join(f"{n} {'text' if n in DICTS else 'double' if t in (abi.kDOUBLE, abi.kFLOAT) else 'bigint'}" for n, t, _, _ in FULL_COLS)
    con.execute(f"CREATE TABLE {name}({decl})")
    con.executemany(f"INSERT INTO {name} VALUES({','.join('?' * len(FULL_COLS))})", rows)
    return con


def translate_strings(rows, plan):
    """Dictionary ids of string-typed targets back to strings (what getNextRow(translate_strings = true) does with the proxy)."""
    cols = []
    for i, t in enumerate(plan.targets[: plan.num_targets]):
        if not t.is_agg and t.sql_type.type in (abi.kTEXT, abi.kVARCHAR, abi.kCHAR):
            cols.append((i, DICTS[FULL_NAMES[t.arg_col_id]]))
    if not cols:
        return rows
    out = []
    for r in rows:
        r = list(r)
        for i, d in cols:
            if r[i] is not None:
                r[i] = d[r[i]]
        out.append(tuple(r))
    return out
