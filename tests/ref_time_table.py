"""The TIME-family / encoded columns of the reference's golden table `test` (Tests/ExecuteTest.cpp:143-186 DDL,
:30063-30115 rows): m timestamp(0), me timestamp(0) encoding fixed(32), n time(0), ne time encoding fixed(32),
o date (DAYS(32), the default), o1 date encoding fixed(16) (= DAYS(16)), o2 date encoding fixed(32) (= DAYS(32)),
fx int encoding fixed(16) — next to x, y, t.  Same three row templates x 10 / 5 / 5 as ref_tables.py."""
import sqlite3

import numpy as np

from heavydb_b200 import abi

DAY = 86400
D_1999_09_09 = 10843 * DAY            # '1999-09-09'
D_1999_09_08 = D_1999_09_09 - DAY     # '1999-09-08', the constant of ExecuteTest.cpp:2040-2048
TS_A, TS_B = 1418509395, 1418595795   # '2014-12-13 22:23:15', '2014-12-14 22:23:15'
T_151314 = 15 * 3600 + 13 * 60 + 14   # '15:13:14'

# name, logical type, notnull, col_encoded_sizes entry
TIME_COLS = [
    ("x", abi.kINT, True, 0), ("y", abi.kINT, False, 0), ("t", abi.kBIGINT, False, 0),
    ("m", abi.kTIMESTAMP, False, 0), ("me", abi.kTIMESTAMP, False, 4),
    ("n", abi.kTIME, False, 0), ("ne", abi.kTIME, False, 4),
    ("o", abi.kDATE, False, -4), ("o1", abi.kDATE, False, -2), ("o2", abi.kDATE, False, -4),
    ("fx", abi.kINT, False, 2),
]
TIME_NAMES = [c[0] for c in TIME_COLS]
_T1 = (7, 42, 1001, TS_A, TS_A, T_151314, T_151314, D_1999_09_09, D_1999_09_09, D_1999_09_09, 9)
_T2 = (8, 43, 1002, TS_A, None, T_151314, None, None, None, None, None)
_T3 = (7, 43, 1002, TS_B, None, T_151314, None, D_1999_09_09, D_1999_09_09, D_1999_09_09, 11)


def time_rows(num_rows: int = 10):
    return [_T1] * num_rows + [_T2] * (num_rows // 2) + [_T3] * (num_rows // 2)


def make_table(rows, fragment_size: int = 2) -> abi.Table:
    t = abi.Table([(ty, nn) for _, ty, nn, _ in TIME_COLS], encoded_sizes=[e for *_, e in TIME_COLS])
    arrays = []
    for c, (_, ty, _nn, enc) in enumerate(TIME_COLS):
        dt = t.physical_dtype(c)
        null = t.physical_null(c)
        if enc < 0:     # days-encoded: the chunk holds days
            arrays.append(np.array([null if r[c] is None else r[c] // DAY for r in rows], dtype=dt))
        else:
            arrays.append(np.array([null if r[c] is None else r[c] for r in rows], dtype=dt))
    for b in range(0, len(rows), fragment_size):
        t.add_host_fragment([a[b:b + fragment_size] for a in arrays])
    return t


def make_sqlite(rows, name="test"):
    con = sqlite3.connect(":memory:")
    con.execute(f"CREATE TABLE {name}({', '.join(n + ' bigint' for n in TIME_NAMES)})")
    con.executemany(f"INSERT INTO {name} VALUES({','.join('?' * len(TIME_NAMES))})", rows)
    return con


def fold_time_literals(sql: str) -> str:
    """'YYYY-MM-DD[ hh:mm:ss]' / 'hh:mm:ss' literals -> the epoch seconds the analyzer folds them to (Constant of type
    DATE / TIMESTAMP / TIME), so that ExecuteTest.cpp query strings can be used verbatim."""
    import datetime as dt
    import re

    def repl(m):
        txt = m.group(1)
        if re.fullmatch(r"\d\d:\d\d:\d\d", txt):
            h, mi, se = map(int, txt.split(":"))
            return str(h * 3600 + mi * 60 + se)
        fmt = "%Y-%m-%d %H:%M:%S" if " " in txt else "%Y-%m-%d"
        return str(int((dt.datetime.strptime(txt, fmt) - dt.datetime(1970, 1, 1)).total_seconds()))
    return re.sub(r"'(\d{4}-\d\d-\d\d(?: \d\d:\d\d:\d\d)?|\d\d:\d\d:\d\d)'", repl, sql)


# verbatim strings of Tests/ExecuteTest.cpp (Select.FilterAndSimpleAggregation :2040-2050, Select.Time :27998)
VERBATIM = [
    "SELECT COUNT(*) FROM test WHERE o1 > '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o1 <= '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o1 = '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o1 <> '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o2 > '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o2 <= '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o2 = '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o2 <> '1999-09-08';",
    "SELECT COUNT(*) FROM test WHERE o1 = o2;",
    "SELECT COUNT(*) FROM test WHERE o1 <> o2;",
]

# ExecuteTest.cpp queries over these columns (date / time literals written as the epoch values the analyzer folds them to)
TIME_QUERIES = [
    f"SELECT COUNT(*) FROM test WHERE o1 > {D_1999_09_08};",       # :2040
    f"SELECT COUNT(*) FROM test WHERE o1 <= {D_1999_09_08};",      # :2041
    f"SELECT COUNT(*) FROM test WHERE o1 = {D_1999_09_08};",       # :2042
    f"SELECT COUNT(*) FROM test WHERE o1 <> {D_1999_09_08};",      # :2043
    f"SELECT COUNT(*) FROM test WHERE o >= {D_1999_09_09};",       # :2044 (CAST('1999-09-09' AS DATE))
    f"SELECT COUNT(*) FROM test WHERE o2 > {D_1999_09_08};",       # :2045
    f"SELECT COUNT(*) FROM test WHERE o2 <= {D_1999_09_08};",      # :2046
    f"SELECT COUNT(*) FROM test WHERE o2 = {D_1999_09_08};",       # :2047
    f"SELECT COUNT(*) FROM test WHERE o2 <> {D_1999_09_08};",      # :2048
    "SELECT COUNT(*) FROM test WHERE o1 = o2;",                    # :2049
    "SELECT COUNT(*) FROM test WHERE o1 <> o2;",                   # :2050
    f"SELECT o, COUNT(*) FROM test WHERE o <= {D_1999_09_09} GROUP BY o ORDER BY 2;",   # :5317
    "SELECT fx, COUNT(*) FROM test GROUP BY fx ORDER BY 2 DESC, 1 ASC NULLS FIRST;",     # :2844 (fx IS NULL DESC == NULLS FIRST)
    f"SELECT COUNT(*) FROM test WHERE m <= {TS_A};",               # :27956 family
    f"SELECT COUNT(*) FROM test WHERE m > {TS_A - 3600};",         # :27998
    "SELECT COUNT(*) FROM test WHERE fx IS NULL;",                 # :5057 as a count
    # the same columns as keys / aggregate arguments
    "SELECT m, COUNT(*), MIN(me), MAX(ne), COUNT(o1) FROM test GROUP BY m;",
    "SELECT o1, o2, COUNT(*), MIN(o), MAX(m) FROM test GROUP BY o1, o2;",
    "SELECT ne, COUNT(*), SUM(y) FROM test WHERE me IS NOT NULL OR n = 54794 GROUP BY ne;",
    "SELECT x, MIN(o2), MAX(o1), COUNT(me), MIN(n), SUM(fx) FROM test GROUP BY x;",
]
