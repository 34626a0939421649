"""Reduction shapes of Tests/GpuSharedMemoryTest.cpp:457-617 and Tests/ResultSetTest.cpp:1026-1105 (tests/reduce_ladder.py):
the oracle keeps one buffer per fragment and folds them with its restatement of ResultSetStorage::reduce — checked
against SQLite, and for independence of the worker-thread count."""
import pytest

import oracle_lib
import reduce_ladder as rl
import ref_tables as rt
import sqlmini


@pytest.mark.parametrize("entry_count", rl.ENTRY_COUNTS)
def test_reduce_ladder(entry_count):
    for i, step in enumerate(rl.STEPS):
        num_buffers = [2, 3, 8, 17, 64, 128][i]
        table, rows = rl.ladder_table(entry_count, step, num_buffers, rows_per_buffer=40, seed=entry_count * 100 + step)
        con = rt.make_sqlite(rows, rl.COLS, "t")
        unit = sqlmini.parse(rl.QUERY, table, rl.NAMES)
        res = oracle_lib.execute(unit, table, num_threads=4)
        rt.assert_rows_match(res.rows(), [tuple(r) for r in con.execute(rl.QUERY.rstrip(";")).fetchall()])
        one = oracle_lib.execute(unit, table, num_threads=1)
        assert bytes(one.buffer()) == bytes(res.buffer())
