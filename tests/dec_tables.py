"""DECIMAL columns of the reference's golden table `test` (Tests/ExecuteTest.cpp:173-174 `dd decimal(10, 2)`,
`dd_notnull decimal(10, 2) not null`; values 111.1 / 222.2 / 333.3 in the three INSERT templates, :30063-30115) plus a
second table with NULLs, negative values and a DECIMAL(7, 3) the DDL stores as 32-bit FIXED chunks.  SQLite holds the
same values as REAL, like the reference's SQLiteComparator does for decimals (ExecuteTest.cpp:467, :595)."""
from __future__ import annotations

import sqlite3

import numpy as np

from heavydb_b200 import abi

# name, sql type, notnull, scale, encoded size
DEC_COLS = [
    ("x", abi.kINT, True, 0, 0),
    ("y", abi.kINT, False, 0, 0),
    ("dd", abi.kDECIMAL, False, 2, 0),
    ("dd_notnull", abi.kDECIMAL, True, 2, 0),
    ("p", abi.kDECIMAL, False, 3, 4),     # DECIMAL(7, 3): precision <= 9 -> ENCODING FIXED(32) chunks
    ("q", abi.kNUMERIC, False, 2, 2),     # NUMERIC(4, 2): precision <= 4 -> FIXED(16)
]
DEC_NAMES = [c[0] for c in DEC_COLS]

# scaled integers (value x 10^scale); None = NULL
_T1 = (7, 42, 11110, 11110, 1500, 125)
_T2 = (8, 43, 22220, 22220, -2500, None)
_T3 = (7, 43, 33330, 33330, None, -999)


def golden_rows(num_rows: int = 10):
    return [_T1] * num_rows + [_T2] * (num_rows // 2) + [_T3] * (num_rows // 2)


def mixed_rows(n: int = 600, seed: int = 11):
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(n):
        dd = None if rng.random() < 0.15 else int(rng.integers(-5000, 5000))
        p = None if rng.random() < 0.2 else int(rng.integers(-9999999, 9999999))
        q = None if rng.random() < 0.2 else int(rng.integers(-9999, 9999))
        rows.append((int(rng.integers(0, 9)), None if rng.random() < 0.1 else int(rng.integers(40, 45)), dd,
                     int(rng.integers(0, 40)) * 25, p, q))
    return rows


def make_table(rows, fragment_size: int = 2) -> abi.Table:
    t = abi.Table([(ty, nn) for _, ty, nn, _, _ in DEC_COLS], encoded_sizes=[e for *_, e in DEC_COLS],
                  col_scales={i: s for i, (_, _, _, s, _) in enumerate(DEC_COLS) if s})
    arrays = []
    for c in range(len(DEC_COLS)):
        dt = t.physical_dtype(c)
        null = t.physical_null(c)
        arrays.append(np.array([null if r[c] is None else r[c] for r in rows], dtype=dt))
    for b in range(0, len(rows), fragment_size):
        t.add_host_fragment([a[b:b + fragment_size] for a in arrays])
    return t


def make_sqlite(rows, name="test"):
    con = sqlite3.connect(":memory:")
    decl = ", ".join(f"{n} {'double' if s else 'bigint'}" for n, _, _, s, _ in DEC_COLS)
    con.execute(f"CREATE TABLE {name}({decl})")
    conv = [tuple(None if v is None else (v / 10 ** DEC_COLS[c][3] if DEC_COLS[c][3] else v) for c, v in enumerate(r)) for r in rows]
    con.executemany(f"INSERT INTO {name} VALUES({','.join('?' * len(DEC_COLS))})", conv)
    return con


# ExecuteTest.cpp:1971-1975, :1984-1986 (the CAST(.. AS decimal(10, 2)) literal is folded by the analyzer), :2823, :12022
GOLDEN_QUERIES = [
    "SELECT MIN(dd) FROM test;",
    "SELECT MAX(dd) FROM test;",
    "SELECT SUM(dd) FROM test;",
    "SELECT AVG(dd) FROM test;",
    "SELECT AVG(dd) FROM test WHERE x > 6 AND x < 8;",
    "SELECT COUNT(*) FROM test WHERE dd > 111.0;",
    "SELECT COUNT(*) FROM test WHERE dd > 222.0;",
    "SELECT COUNT(*) FROM test WHERE dd > 333.0;",
    "SELECT x, dd, COUNT(*) FROM test GROUP BY x, dd ORDER BY x, dd;",
    "SELECT SUM(dd) FROM test WHERE x > 8;",
]
MORE_QUERIES = [
    "SELECT dd, COUNT(*), SUM(dd_notnull), AVG(p) FROM test GROUP BY dd;",
    "SELECT dd_notnull, MIN(dd), MAX(dd), COUNT(dd) FROM test GROUP BY dd_notnull;",
    "SELECT x, SUM(dd), AVG(dd), MIN(p), MAX(p), SUM(q), AVG(q) FROM test GROUP BY x;",
    "SELECT COUNT(*), SUM(p), AVG(p), MIN(q), MAX(q) FROM test WHERE dd BETWEEN -10.5 AND 20.25;",
    "SELECT COUNT(*) FROM test WHERE p >= -1234.567 AND p < 5000;",
    "SELECT COUNT(*), AVG(dd) FROM test WHERE q IN (1.25, -9.99, 3.5) OR dd IS NULL;",
    "SELECT COUNT(*) FROM test WHERE dd = dd_notnull;",
    "SELECT COUNT(*) FROM test WHERE dd <> 1.5 AND p IS NOT NULL;",
    "SELECT y, AVG(dd_notnull), SUM(p) FROM test WHERE dd_notnull <= 5 GROUP BY y ORDER BY 2 DESC, 1;",
    "SELECT q, COUNT(*) FROM test WHERE q > 0 GROUP BY q ORDER BY 2 DESC, 1 LIMIT 7;",
    "SELECT dd_notnull, SUM(dd) FROM test GROUP BY dd_notnull ORDER BY 2 LIMIT 5 OFFSET 2;",
    "SELECT MIN(dd), MAX(dd), SUM(dd), AVG(dd), COUNT(dd) FROM test WHERE x > 100;",
]


FACT_NAMES, DIM_NAMES = ["fk", "v"], ["id", "price"]
JOIN_QUERIES = [
    "SELECT d.price, COUNT(*), SUM(t.v) FROM t JOIN d ON t.fk = d.id GROUP BY d.price;",
    "SELECT COUNT(*), SUM(d.price), AVG(d.price), MIN(d.price), MAX(d.price), COUNT(d.price) FROM t JOIN d ON t.fk = d.id WHERE d.price > 2.5;",
    "SELECT COUNT(*), SUM(d.price), AVG(d.price) FROM t LEFT JOIN d ON t.fk = d.id WHERE d.price <= 7.25 OR d.price IS NULL;",
    "SELECT d.price, COUNT(*), AVG(t.v) FROM t LEFT JOIN d ON t.fk = d.id WHERE t.v < 50 GROUP BY d.price;",
]


def star_join(seed: int = 8):
    """fact(fk INT, v BIGINT NOT NULL) x dim(id INT NOT NULL, price DECIMAL(7, 2) as FIXED(32), three NULL prices)."""
    rng = np.random.default_rng(seed)
    dim = abi.Table([(abi.kINT, True), (abi.kDECIMAL, False)], encoded_sizes=[0, 4], col_scales={1: 2})
    price = rng.integers(0, 60, 50).astype(np.int32) * 25           # multiples of 0.25: exact as REAL
    price[[3, 17, 40]] = -2**31                                      # NULL (the FIXED(32) sentinel)
    dim.add_host_fragment([np.arange(50, dtype=np.int32), price])
    fact = abi.Table([(abi.kINT, False), (abi.kBIGINT, True)])
    fk_all, v_all = [], []
    for _ in range(3):
        fk = rng.integers(-3, 55, 800).astype(np.int32)
        fk[rng.random(800) < 0.05] = abi.NULL_INT
        v = rng.integers(0, 100, 800).astype(np.int64)
        fact.add_host_fragment([fk, v])
        fk_all += fk.tolist()
        v_all += v.tolist()
    return fact, dim, price, fk_all, v_all
