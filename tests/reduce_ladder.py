"""The shapes of the reference's reduction tests restated as queries: Tests/GpuSharedMemoryTest.cpp:457-617 reduces
`num_buffers` partial tables of `entry_count` entries (entry counts {1,2,3,5,13,31,63,126,241,511,1021}, steps
{2,3,5,7,11,13}, buffers {2..128}) and compares with the CPU reduce; Tests/ResultSetTest.cpp:1026-1105 reduces two
storages whose entries interleave (every `step`-th entry filled).  Here: `entry_count` keys, one fragment per partial
buffer, fragment f holding only the keys k with k % step == f % step (so partial tables interleave and overlap)."""
import numpy as np

from heavydb_b200 import abi

ENTRY_COUNTS = [1, 2, 3, 5, 13, 31, 63, 126, 241, 511, 1021]
STEPS = [2, 3, 5, 7, 11, 13]
COLS = [("k", abi.kINT, True), ("v", abi.kBIGINT, False), ("d", abi.kDOUBLE, False), ("w", abi.kSMALLINT, True)]
NAMES = [c[0] for c in COLS]
QUERY = "SELECT k, COUNT(*), SUM(v), MIN(v), MAX(w), AVG(d), COUNT(v) FROM t GROUP BY k;"


def ladder_table(entry_count, step, num_buffers, rows_per_buffer, seed):
    rng = np.random.default_rng(seed)
    t = abi.Table([(ty, nn) for _, ty, nn in COLS])
    rows = []
    for f in range(num_buffers):
        keys = np.arange(entry_count, dtype=np.int32)
        keys = keys[keys % step == f % step]
        if keys.size == 0:
            keys = np.array([0], dtype=np.int32)
        k = rng.choice(keys, rows_per_buffer).astype(np.int32)
        v = rng.integers(-10**6, 10**6, rows_per_buffer).astype(np.int64)
        v[rng.random(rows_per_buffer) < 0.15] = abi.NULL_BIGINT
        d = rng.normal(0, 100, rows_per_buffer)
        d[rng.random(rows_per_buffer) < 0.15] = abi.NULL_DOUBLE
        w = rng.integers(-3000, 3000, rows_per_buffer).astype(np.int16)
        t.add_host_fragment([k, v, d, w])
        for i in range(rows_per_buffer):
            rows.append((int(k[i]), None if v[i] == abi.NULL_BIGINT else int(v[i]), None if d[i] == abi.NULL_DOUBLE else float(d[i]), int(w[i])))
    return t, rows
