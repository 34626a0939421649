"""The cardinality-estimation query of the baseline-hash flow (SURVEY §8b "errors": CardinalityEstimationRequired ->
RelAlgExecutor::getNDVEstimation -> executeWorkUnit with an NDVEstimator, CardinalityEstimator.cpp:54-116).  CPU side:
the oracle's linear_probabilistic_count / getNDVEstimator restatement against exact distinct counts, and planner parity."""
import numpy as np
import pytest

import join_tables as jt
import oracle_lib
from heavydb_b200 import abi, executor
from test_gpu_parity import RAND_NAMES, random_table

EST_CASES = [["sparse"], ["k32"], ["k16", "nn32"], ["big"], ["k8", "k64", "nn64"], ["a8"]]


def estimator_unit(table, names, cols, large=False, inner=None, filt=None):
    b = abi.UnitBuilder(table)
    if inner is not None:
        b.join(inner[0], inner[1], inner[2])
    if filt is not None:
        b.add_qual(b.cmp(names.index(filt[0]), filt[1], filt[2]), simple=True)
    b.estimator([c if isinstance(c, tuple) else names.index(c) for c in cols], large=large)
    return b.build()


@pytest.fixture(scope="module")
def table():
    return random_table(60000, seed=4, frag_rows=13000)


@pytest.mark.parametrize("cols", EST_CASES)
def test_oracle_estimate_is_close_and_plans_agree(table, cols):
    unit = estimator_unit(table, RAND_NAMES, cols)
    res = oracle_lib.execute(unit, table, num_threads=4)
    arrays = [np.concatenate([f.host_cols[RAND_NAMES.index(c)] for f in table.fragments]) for c in cols]
    exact = len(set(zip(*[a.tolist() for a in arrays])))
    est = res.ndv_estimator()
    assert abs(est - exact) <= max(3, 0.01 * exact), (est, exact)     # linear counting with 8 Mi bits: < 1 % here
    p = res.plan
    assert p.query_desc_type == abi.Estimator and p.buffer_size == 1 << 20 and p.entry_count == 1
    assert executor.Executor().plan(unit, table).as_dict() == oracle_lib.plan(unit, table).as_dict()
    assert res.rows() == [] and res.row_count() == 0


def test_known_values():
    """getNDVEstimator: 1 for an empty bitmap (CardinalityEstimator.cpp:37-40), -m ln(unset / m) otherwise; one distinct
    tuple sets exactly one bit whatever the row count."""
    t = abi.Table([(abi.kBIGINT, True)])
    t.add_host_fragment([np.full(1000, 12345, dtype=np.int64)])
    res = oracle_lib.execute(estimator_unit(t, ["k"], ["k"]), t)
    buf = res.buffer().view(np.uint8)
    assert int(np.unpackbits(buf).sum()) == 1
    bit_pos = 342635441 % (8 << 20)          # MurmurHash3(&int64{12345}, 8, 0) = 342635441 (SURVEY.md §8c probe value)
    assert buf.view(np.uint32)[bit_pos // 32] == 1 << (bit_pos % 32)
    assert res.ndv_estimator() == int(-(8 << 20) * np.log(1 - 1 / (8 << 20)))
    empty = abi.Table([(abi.kBIGINT, True)])
    empty.add_host_fragment([np.zeros(0, dtype=np.int64)])
    assert oracle_lib.execute(estimator_unit(empty, ["k"], ["k"]), empty).ndv_estimator() == 1


def test_estimator_with_filter_and_join(table):
    unit = estimator_unit(table, RAND_NAMES, ["k32"], filt=("nn32", abi.kLT, 100))
    res = oracle_lib.execute(unit, table)
    k32 = np.concatenate([f.host_cols[RAND_NAMES.index("k32")] for f in table.fragments])
    nn32 = np.concatenate([f.host_cols[RAND_NAMES.index("nn32")] for f in table.fragments])
    exact = len(set(k32[nn32 < 100].tolist()))
    assert abs(res.ndv_estimator() - exact) <= 3
    fact, dim = jt.fact_table(20000, seed=2, frag_rows=6000), jt.dim_table()
    unit = estimator_unit(fact, jt.FACT_NAMES, [(jt.DIM_NAMES.index("big"), 1), "x"], inner=(dim, 0, 0))
    res = oracle_lib.execute(unit, fact)
    assert executor.Executor().plan(unit, fact).as_dict() == oracle_lib.plan(unit, fact).as_dict()
    assert 0.9 * 20000 * 0.8 < res.ndv_estimator() <= 20000      # (dimension row, x) pairs of the matching rows


def test_bad_estimator_units(table):
    b = abi.UnitBuilder(table)
    b.group_by(0)
    b.estimator([0])
    for f in (oracle_lib.plan, executor.Executor().plan):
        with pytest.raises((oracle_lib.OracleError, executor.QueryExecutionError)):
            f(b.build(), table)
    b = abi.UnitBuilder(table)
    b.estimator([RAND_NAMES.index("d")])
    with pytest.raises(oracle_lib.OracleError):
        oracle_lib.plan(b.build(), table)
    with pytest.raises(executor.UnsupportedOnThisPath):
        executor.Executor().plan(b.build(), table)
