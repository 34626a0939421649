"""The reference's storage-level known-answer vectors, replayed: Tests/ResultSetTest.cpp `test_iterate` (:879-930) and
`test_reduce` (:1026-1111) fill result-set STORAGE by hand with the number generators of ResultSetTestUtils.h:33-70
(`fill_storage_buffer_*`, ResultSetTestUtils.cpp:251-466: entry i holds the generator's next value v when i % step == 0 — its
key, every target slot, COUNT slots too, AVG as the pair (v, 1) — and is empty otherwise), then read it back
(`getNextRow`: every target == v) or reduce two storages with ResultSetManager::reduce and read the result
(SUM / COUNT == step * row, everything else == row).

Here the storage descriptors are the ones this path's planner produces for a query (keyed and keyless perfect hash, 8- and
4-byte slots, row-wise and columnar, baseline hash) — the reference's hand-built `generate_test_target_infos` list holds a
projected DOUBLE and a dictionary string that no aggregate query of this path yields — and three readers have to agree with
the expected values: the oracle's restatement of ResultSetIteration, the oracle's restatement of ResultSetStorage::reduce /
ResultSetManager::reduce (incl. the growth + moveEntriesToBuffer of the baseline case), and the PRODUCT's read-out over the
same bytes (`b2q_rs_create_from_storage`, host only)."""

import numpy as np
import pytest

import oracle_lib
import sqlmini
from heavydb_b200 import abi, executor

EMPTY_KEY_64 = np.iinfo(np.int64).max
EMPTY_KEY_32 = np.iinfo(np.int32).max
NAMES = ["k", "k4", "a", "an", "d", "s"]


def make_table(n=100):
    """Only the chunk stats matter (they fix the descriptor: key range [0, n-1], NOT NULL): the storage is filled by hand."""
    t = abi.Table([(abi.kBIGINT, True), (abi.kINT, True), (abi.kINT, True), (abi.kINT, False), (abi.kDOUBLE, True), (abi.kBIGINT, True)])
    k = np.arange(n, dtype=np.int64)
    an = (k % 7).astype(np.int32)
    an[3] = abi.NULL_INT
    sparse = k * 7919 * 10**9          # range far beyond any perfect-hash table: baseline hash
    t.add_host_fragment([k, k.astype(np.int32), (k % 5).astype(np.int32), an, k.astype(np.float64), sparse])
    return t


class Even:                              # EvenNumberGenerator
    def __init__(self):
        self.reset()

    def reset(self):
        self.crt = 0

    def next(self):
        v = self.crt
        self.crt += 2
        return v


class ReverseOddOrEven:                  # ReverseOddOrEvenNumberGenerator(init)
    def __init__(self, init):
        self.init = init
        self.reset()

    def reset(self):
        self.crt = self.init

    def next(self):
        v = self.crt
        self.crt -= 2
        return v


def write_slot(buf, off, width, value):
    buf[off:off + width] = np.array([value], dtype=np.int64 if width == 8 else np.int32).view(np.uint8)


def slot_off(plan, e, s):
    return plan.slot_offset[s] + e * plan.slot_padded_width[s] if plan.output_columnar else e * plan.row_size + plan.slot_offset[s]


def key_off(plan, e):
    return e * 8 if plan.output_columnar else e * plan.row_size


def fill_entry(buf, plan, e, v):
    """fill_one_entry_no_collisions / fill_one_entry_one_col / fill_one_entry_baseline: v into every slot in the target's type,
    AVG as (v, 1)."""
    for t in plan.targets[: plan.num_targets]:
        s = t.first_slot
        w = plan.slot_padded_width[s]
        if w == 0:
            continue                     # a baseline key target reads the key
        fp = t.sql_type.type == abi.kDOUBLE and not (t.is_agg and t.agg_kind == abi.kAVG and t.agg_arg_type.type != abi.kDOUBLE)
        if fp:
            buf[slot_off(plan, e, s):slot_off(plan, e, s) + 8] = np.array([float(v)], dtype=np.float64).view(np.uint8)
        else:
            write_slot(buf, slot_off(plan, e, s), w, v)
        if t.is_agg and t.agg_kind == abi.kAVG:
            write_slot(buf, slot_off(plan, e, s + 1), plan.slot_padded_width[s + 1], 1)


def fill_perfect(plan, gen, step, empties_at_init=False):
    """fill_storage_buffer_perfect_hash_rowwise / _colwise.  Empty entries: EMPTY_KEY and 0xdeadbeef in a keyed layout; in a
    keyless one the reference writes 0 because its test descriptors carry no init values — here the planned init values,
    which is what `isEmptyEntry` of a planned descriptor compares the marker slot with."""
    buf = np.zeros(plan.buffer_size, dtype=np.uint8)
    keyed = not plan.keyless_hash
    kw = 8 if plan.output_columnar else plan.effective_key_width
    gen.reset()
    for e in range(plan.entry_count):
        if e % step == 0:
            v = gen.next()
            if keyed:
                write_slot(buf, key_off(plan, e), kw, v)
            fill_entry(buf, plan, e, v)
        else:
            if keyed:
                write_slot(buf, key_off(plan, e), kw, EMPTY_KEY_64 if kw == 8 else EMPTY_KEY_32)
            for s in range(plan.num_slots):
                w = plan.slot_padded_width[s]
                if w:
                    write_slot(buf, slot_off(plan, e, s), w, 0xdeadbeef if keyed and not empties_at_init else (plan.init_vals[s] if w == 8 else np.int64(plan.init_vals[s]).astype(np.int32)))
    return buf


def fill_baseline(plan, gen, step):
    """fill_storage_buffer_baseline_rowwise: every entry empty, then the keys inserted with get_group_value (the oracle's, pinned
    against the reference's compiled GroupByRuntime.cpp in test_oracle_ref.py)."""
    assert not plan.output_columnar and plan.effective_key_width == 8 and plan.row_size % 8 == 0
    buf = np.zeros(plan.buffer_size, dtype=np.uint8)
    for e in range(plan.entry_count):
        write_slot(buf, key_off(plan, e), 8, EMPTY_KEY_64)
        for s in range(plan.num_slots):
            if plan.slot_padded_width[s]:
                write_slot(buf, slot_off(plan, e, s), plan.slot_padded_width[s], 0 if plan.targets_kind[s] == abi.kCOUNT else 0xdeadbeef)
    gen.reset()
    L = oracle_lib.lib()
    for _ in range(0, plan.entry_count, step):
        v = gen.next()
        key = np.array([v], dtype=np.int64)
        off = L.oracle_get_group_value(buf.ctypes.data, plan.entry_count, key.ctypes.data, 1, 8, plan.row_size // 8)
        assert off >= 0
        e = (off * 8 - 8) // plan.row_size
        assert np.frombuffer(buf[key_off(plan, e):key_off(plan, e) + 8].tobytes(), dtype=np.int64)[0] == v
        fill_entry(buf, plan, e, v)
    return buf


def with_slot_kinds(plan):
    kinds = [None] * plan.num_slots
    for t in plan.targets[: plan.num_targets]:
        kinds[t.first_slot] = t.agg_kind if t.is_agg else None
        if t.is_agg and t.agg_kind == abi.kAVG:
            kinds[t.first_slot + 1] = abi.kCOUNT
    plan.targets_kind = kinds
    return plan


def expected_row(plan, v, count_sum_factor):
    """test_iterate: every target reads v.  test_reduce: SUM / COUNT read factor * v, the others v."""
    row = []
    for t in plan.targets[: plan.num_targets]:
        if t.is_agg and t.agg_kind in (abi.kSUM, abi.kCOUNT):
            x = count_sum_factor * v
            row.append(float(x) if t.sql_type.type == abi.kDOUBLE else x)
        elif t.is_agg and t.agg_kind == abi.kAVG:
            row.append(float(v))
        else:
            row.append(float(v) if t.sql_type.type == abi.kDOUBLE else v)
    return tuple(row)


def is_empty_by_marker(plan, v):
    """A keyless entry whose marker slot holds the marker's init value reads as empty (ResultSetStorage::isEmptyEntry): the
    generator's v == 0 in a COUNT marker, exactly as in the reference's keyless runs of these tests."""
    if not plan.keyless_hash:
        return False
    for t in plan.targets[: plan.num_targets]:
        if t.first_slot == plan.idx_target_as_key:
            return v == plan.init_vals[t.first_slot]
    return False


# (sql, columnar): keyed 8-byte slots; keyed; keyless 8-byte (COUNT marker); keyless 4-byte compact slots; 4-byte key column
PERFECT = [
    ("SELECT k, MIN(an), MAX(an), SUM(an) FROM t GROUP BY k;", False),
    ("SELECT k, MIN(an), MAX(an), SUM(an) FROM t GROUP BY k;", True),
    ("SELECT k, COUNT(*), SUM(a), AVG(a), MIN(d), MAX(a) FROM t GROUP BY k;", False),
    ("SELECT k, COUNT(*), SUM(a), AVG(a), MIN(d), MAX(a) FROM t GROUP BY k;", True),
    ("SELECT k4, COUNT(*), MIN(a) FROM t GROUP BY k4;", False),
    ("SELECT k4, COUNT(*), MIN(a) FROM t GROUP BY k4;", True),
    ("SELECT k4, MIN(an), SUM(an) FROM t GROUP BY k4;", False),
    ("SELECT MIN(an), k, SUM(an), AVG(an) FROM t GROUP BY k;", False),
    ("SELECT k4, COUNT(*) FROM t GROUP BY k4;", False),           # pick_target_compact_width = 4: 4-byte slots
    ("SELECT k4, COUNT(*) FROM t GROUP BY k4;", True),
    ("SELECT COUNT(*), k4, COUNT(a) FROM t GROUP BY k4;", False),
]


def product_rows(unit, table, storage, columnar=False, **kw):
    eo = executor.execution_options(output_columnar_hint=columnar)
    rs = executor.Executor().resultSetFromStorage(storage, unit, table, eo=eo, **kw)
    return rs.rows(), rs


@pytest.mark.parametrize("sql,columnar", PERFECT)
def test_iterate_perfect_hash(sql, columnar):
    """TEST(Iterate, PerfectHashOneCol*) (ResultSetTest.cpp:1398-1470): EvenNumberGenerator, step 2."""
    table = make_table()
    unit = sqlmini.parse(sql, table, NAMES)
    plan = with_slot_kinds(oracle_lib.plan(unit, table, output_columnar=columnar))
    assert plan.query_desc_type == abi.GroupByPerfectHash and plan.entry_count == 100
    storage = fill_perfect(plan, Even(), 2)
    want = [expected_row(plan, v, 1) for v in range(0, 100, 2) if not is_empty_by_marker(plan, v)]
    ours = oracle_lib.result_from_storage(unit, table, storage, output_columnar=columnar)
    assert ours.rows() == want
    got, rs = product_rows(unit, table, storage, columnar)
    assert got == want
    assert rs.rowCount() == len(want) and rs.entryCount() == 100
    for e in range(100):
        assert rs.isRowAtEmpty(e) == (e % 2 == 1 or is_empty_by_marker(plan, e))


@pytest.mark.parametrize("sql,columnar", PERFECT)
def test_reduce_perfect_hash(sql, columnar):
    """TEST(Reduce, PerfectHashOneCol*) (ResultSetTest.cpp:1591-1700): two storages from two EvenNumberGenerators, step 2:
    SUM / COUNT == step * row_idx, the others == row_idx."""
    table = make_table()
    unit = sqlmini.parse(sql, table, NAMES)
    plan = with_slot_kinds(oracle_lib.plan(unit, table, output_columnar=columnar))
    step = 2
    s1, s2 = fill_perfect(plan, Even(), step), fill_perfect(plan, Even(), step)
    r1 = oracle_lib.result_from_storage(unit, table, s1, output_columnar=columnar)
    r2 = oracle_lib.result_from_storage(unit, table, s2, output_columnar=columnar)
    red = oracle_lib.reduce_result_sets([r1, r2])
    want = [expected_row(plan, v, step) for v in range(0, 100, step) if not is_empty_by_marker(plan, v)]
    assert red.rows() == want
    got, rs = product_rows(unit, table, red.buffer(), columnar)         # the product's read-out of the reduced bytes
    assert got == want
    # ... and entry by entry, as test_reduce reads it: getRowAtNoTranslations(row_idx), an empty row for an empty entry
    for row_idx in range(100):
        row = rs.getRowAt(row_idx)
        if row_idx % step or is_empty_by_marker(plan, row_idx):
            assert row == () and rs.isRowAtEmpty(row_idx)
        else:
            assert row == expected_row(plan, row_idx, step)
    assert rs.getRowAt(100) == () and rs.getRowAt(10**9) == ()
    # three storages, as reduceMultiDeviceResultSets folds one per device: COUNT / SUM grow by v per storage
    red3 = oracle_lib.reduce_result_sets([r1, r2, oracle_lib.result_from_storage(unit, table, s1, output_columnar=columnar)])
    assert red3.rows() == [expected_row(plan, v, 3) for v in range(0, 100, step) if not is_empty_by_marker(plan, v)]


def test_reduce_perfect_hash_disjoint_entries():
    """Storages whose non-empty entries do not overlap: the reduce copies the key and folds the values of `that` into entries
    `this` never touched (reduceOneEntryNoCollisions :398-450) — which works because an untouched entry of a REAL buffer holds
    the init values (the 0xdeadbeef of the reference's test filler only survives there because both of its storages fill the
    same entries)."""
    table = make_table()
    unit = sqlmini.parse("SELECT k, MIN(an), MAX(an), SUM(an) FROM t GROUP BY k;", table, NAMES)
    plan = with_slot_kinds(oracle_lib.plan(unit, table))
    s1 = fill_perfect(plan, Even(), 2, empties_at_init=True)   # entries 0, 2, 4, ... hold 0, 2, 4, ...
    s2 = np.array(s1, copy=True)
    # second storage: entry e holds value e for the ODD entries, even ones empty
    for e in range(100):
        if e % 2:
            write_slot(s2, key_off(plan, e), 8, e)
            fill_entry(s2, plan, e, e)
        else:
            write_slot(s2, key_off(plan, e), 8, EMPTY_KEY_64)
    red = oracle_lib.reduce_result_sets([oracle_lib.result_from_storage(unit, table, s1), oracle_lib.result_from_storage(unit, table, s2)])
    want = [expected_row(plan, v, 1) for v in range(100)]
    assert red.rows() == want
    assert product_rows(unit, table, red.buffer())[0] == want


@pytest.mark.parametrize("n", [4, 37])
def test_reduce_baseline_hash(n):
    """TEST(Reduce, BaselineHash) (ResultSetTest.cpp:1811-1817): EvenNumberGenerator vs ReverseOddOrEvenNumberGenerator(2n - 1),
    step 1, two FULL n-entry tables; ResultSetManager::reduce builds a 2n-entry storage, moves the first set's entries into it
    and re-probes the second's; sorted by the first column, row r reads r in every target (step * row_idx with step 1)."""
    table = make_table()
    unit = sqlmini.parse("SELECT s, COUNT(*), SUM(a), MIN(d), AVG(a) FROM t GROUP BY s;", table, NAMES)
    plan = with_slot_kinds(oracle_lib.plan(unit, table, entry_guess=n, has_card=True))
    assert plan.query_desc_type == abi.GroupByBaselineHash and plan.entry_count == n
    s1 = fill_baseline(plan, Even(), 1)
    s2 = fill_baseline(plan, ReverseOddOrEven(2 * n - 1), 1)
    r1 = oracle_lib.result_from_storage(unit, table, s1, entry_guess=n, has_card=True)
    r2 = oracle_lib.result_from_storage(unit, table, s2, entry_guess=n, has_card=True)
    assert sorted(r1.rows()) == [expected_row(plan, v, 1) for v in range(0, 2 * n, 2)]
    assert sorted(product_rows(unit, table, s2, max_groups_buffer_entry_guess=n, has_cardinality_estimation=True)[0]) == \
        [expected_row(plan, v, 1) for v in range(1, 2 * n, 2)]
    red = oracle_lib.reduce_result_sets([r1, r2])
    assert red.entry_count() == 2 * n
    want = [expected_row(plan, v, 1) for v in range(2 * n)]
    assert sorted(red.rows()) == want
    red.sort([(1, False, False)])                            # the test's own `sort` flag: ORDER BY the first column
    assert red.rows() == want
    got, rs = product_rows(unit, table, red.buffer(), max_groups_buffer_entry_guess=2 * n, has_cardinality_estimation=True)
    assert sorted(got) == want and rs.entryCount() == 2 * n
    # the same key in both storages: the re-probe finds the moved entry and reduces into it
    both = oracle_lib.reduce_result_sets([r1, oracle_lib.result_from_storage(unit, table, s1, entry_guess=n, has_card=True)])
    assert sorted(both.rows()) == [expected_row(plan, v, 2) for v in range(0, 2 * n, 2)]
