"""Tables with `ENCODING FIXED(bits)` columns and a $deleted$ column (the step before the path —
what real HeavyDB chunks look like: FixedLengthEncoder chunks, DataMgr/Encoder.h; deleted-rows column,
Execute.cpp:4593)."""
import numpy as np

from heavydb_b200 import abi

ENC_COLS = [
    # name, logical type, notnull, encoded physical bytes (0 = none)
    ("k_i32_f16", abi.kINT, False, 2),       # INT ENCODING FIXED(16), nullable  — like `fx` in ExecuteTest.cpp:169
    ("k_i64_f32", abi.kBIGINT, True, 4),     # BIGINT NOT NULL ENCODING FIXED(32)
    ("a_i64_f8", abi.kBIGINT, False, 1),     # BIGINT ENCODING FIXED(8), nullable
    ("a_i32_f8", abi.kINT, False, 1),
    ("a_i64_f16", abi.kBIGINT, True, 2),
    ("plain64", abi.kBIGINT, False, 0),
    ("d", abi.kDOUBLE, True, 0),
    ("deleted", abi.kBOOLEAN, True, 0),
]
ENC_NAMES = [c[0] for c in ENC_COLS]


def enc_table(n, seed, frag_rows, deleted_frac=0.2, fully_deleted_fragment=None, with_deleted=True):
    rng = np.random.default_rng(seed)

    def nulls(a, phys_null, p=0.15):
        a = a.copy()
        if n:
            a[rng.random(n) < p] = phys_null
        return a
    cols = [
        nulls(rng.integers(-40, 60, n).astype(np.int16), -2**15),
        rng.integers(1000, 1400, n).astype(np.int32),
        nulls(rng.integers(-100, 100, n).astype(np.int8), -2**7),
        nulls(rng.integers(-127, 128, n).astype(np.int8), -2**7),
        rng.integers(-30000, 30000, n).astype(np.int16),
        nulls(rng.integers(-2**40, 2**40, n).astype(np.int64), -2**63),
        rng.random(n),
        (rng.random(n) < deleted_frac).astype(np.int8),
    ]
    if fully_deleted_fragment is not None:
        b = fully_deleted_fragment * frag_rows
        cols[7][b:b + frag_rows] = 1
    t = abi.Table([(ty, nn) for _, ty, nn, _ in ENC_COLS], encoded_sizes=[e for *_, e in ENC_COLS],
                  deleted_column=7 if with_deleted else None)
    for b in range(0, max(n, 1), frag_rows):
        t.add_host_fragment([c[b:b + frag_rows] for c in cols])
    return t


ENC_QUERIES = [
    "SELECT COUNT(*), COUNT(a_i64_f8), SUM(a_i64_f8), MIN(a_i64_f8), MAX(a_i64_f8), AVG(a_i64_f8) FROM e;",
    "SELECT k_i32_f16, COUNT(*), SUM(a_i32_f8), MIN(a_i32_f8), MAX(a_i64_f16), AVG(a_i64_f16) FROM e GROUP BY k_i32_f16;",
    "SELECT k_i64_f32, SUM(plain64), COUNT(a_i64_f8), AVG(d) FROM e WHERE a_i64_f16 > -1000 GROUP BY k_i64_f32;",
    "SELECT k_i32_f16, MIN(k_i32_f16), MAX(k_i32_f16), SUM(k_i32_f16) FROM e WHERE k_i32_f16 <> 7 GROUP BY k_i32_f16;",
    "SELECT COUNT(*) FROM e WHERE a_i32_f8 >= -128 AND a_i64_f8 < 50;",
    "SELECT COUNT(*), SUM(a_i64_f16) FROM e WHERE k_i32_f16 < 10 OR a_i32_f8 = 5 OR d > 0.9;",
    "SELECT a_i64_f8, COUNT(*) FROM e GROUP BY a_i64_f8;",
    "SELECT plain64, COUNT(*), SUM(a_i32_f8) FROM e GROUP BY plain64;",            # baseline hash
    "SELECT k_i64_f32, COUNT(*) FROM e WHERE k_i64_f32 >= 1100 AND k_i64_f32 <= 1200 GROUP BY k_i64_f32;",
]
