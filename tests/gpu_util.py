"""Helpers for the `-m gpu` parity tests: run a query through the C ABI (CUDA path) and through the oracle on the
same inputs and compare rows + raw output buffers."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

import oracle_lib
import ref_tables
from heavydb_b200 import abi, executor

FP_RTOL = 1e-6  # north_star: SUM/AVG(double) within 1e-6 relative; everything integer is bit-exact
# SUM / AVG of a FLOAT column: the reference accumulates in float precision in whatever order its threads arrive
# (agg_sum_float, atomicAdd(float) on its GPU); the product accumulates the exactly-widened values in double and rounds
# once.  The two agree to float rounding of the partial sums, not to 1e-6; sums that cancel get the same fraction of a
# typical partial sum as an absolute bound (ref_tables.float_sum_atol).
FLOAT_SUM_RTOL = ref_tables.FLOAT_SUM_RTOL


def _cudart():
    """Device memory for tests comes from torch (plumbing only)."""
    import torch
    return torch


def has_gpu() -> bool:
    try:
        return executor.lib().b2q_device_count() > 0
    except Exception:
        return False


class DeviceTable:
    """Copies a host abi.Table to the GPU (torch tensors own the memory) and exposes it as a GPU_LEVEL table."""

    def __init__(self, table: abi.Table):
        torch = _cudart()
        self.keep = []
        self.table = abi.Table(table.col_types, encoded_sizes=table.encoded_sizes, deleted_column=table.deleted_column, col_scales=table.col_scales)
        for f in table.fragments:
            ptrs = []
            for a in f.host_cols:
                if a is None or a.size == 0:
                    ptrs.append(0)
                    continue
                t = torch.from_numpy(a.view(np.uint8).copy()).cuda()
                self.keep.append(t)
                ptrs.append(t.data_ptr())
            self.table.add_device_fragment(f.num_tuples, ptrs, f.stats, fragment_id=f.fragment_id)
        torch.cuda.synchronize()


def column_tolerances(plan, n_rows):
    """(rtol, atol) per output column: 1e-6 relative, except SUM / AVG over a FLOAT argument."""
    out = []
    for t in plan.targets[: plan.num_targets]:
        if t.is_agg and t.agg_kind in (abi.kSUM, abi.kAVG) and t.agg_arg_type.type == abi.kFLOAT:
            out.append((FLOAT_SUM_RTOL, ref_tables.float_sum_atol(n_rows)))
        else:
            out.append((FP_RTOL, 0.0))
    return out


def rows_equal(got, want, fp_rtol=FP_RTOL, col_tol=None):
    def key(r):
        return tuple((0, 0) if v is None else (1, v) for v in r)
    g, w = sorted(got, key=key), sorted(want, key=key)
    assert len(g) == len(w), f"row count {len(g)} != {len(w)}\n got={g[:8]}\nwant={w[:8]}"
    for a, b in zip(g, w):
        assert len(a) == len(b)
        for c, (va, vb) in enumerate(zip(a, b)):
            rtol, atol = col_tol[c] if col_tol else (fp_rtol, 0.0)
            if va is None or vb is None:
                assert va is None and vb is None, f"NULL mismatch: {a} vs {b}"
            elif isinstance(vb, float):
                assert isinstance(va, float)
                if math.isnan(vb):
                    assert math.isnan(va)
                else:
                    assert va == vb or abs(va - vb) <= rtol * abs(vb) + atol, f"{a} vs {b}"
            else:
                assert va == vb, f"{a} vs {b}"


def rows_equal_ordered(got, want, fp_rtol=FP_RTOL):
    """Row by row, in order (ORDER BY results)."""
    assert len(got) == len(want), f"row count {len(got)} != {len(want)}\n got={got[:8]}\nwant={want[:8]}"
    for i, (a, b) in enumerate(zip(got, want)):
        assert len(a) == len(b)
        for va, vb in zip(a, b):
            if va is None or vb is None:
                assert va is None and vb is None, f"row {i}: NULL mismatch: {a} vs {b}"
            elif isinstance(vb, float):
                assert isinstance(va, float) and (va == vb or abs(va - vb) <= fp_rtol * abs(vb)), f"row {i}: {a} vs {b}"
            else:
                assert va == vb, f"row {i}: {a} vs {b}"


def _entry_bytes(buf, plan, entries):
    """Per-entry byte image [keys | slots] of the given entries of a row-wise or columnar buffer, plus the byte
    ranges (inside that image) of floating-point SUM slots."""
    b = buf.view(np.int8)
    n = plan.entry_count
    entries = np.asarray(entries, dtype=np.int64)
    fp_sum = {t.first_slot for t in plan.targets[: plan.num_targets]
              if t.is_agg and t.agg_kind in (abi.kSUM, abi.kAVG) and t.agg_arg_type.type == abi.kDOUBLE}
    parts, fp_ranges, off = [], [], 0
    keyed = plan.query_desc_type != abi.NonGroupedAggregate and not plan.keyless_hash
    if plan.output_columnar:
        if keyed:
            stride = (8 * n + 7) // 8 * 8
            for c in range(max(plan.num_group_cols, 1)):
                parts.append(b[c * stride:c * stride + 8 * n].reshape(n, 8)[entries]); off += 8
        for s in range(plan.num_slots):
            w = plan.slot_padded_width[s]
            if not w:
                continue
            o = plan.slot_offset[s]
            parts.append(b[o:o + w * n].reshape(n, w)[entries])
            if s in fp_sum:
                fp_ranges.append((off, off + 8))
            off += w
        return (np.concatenate(parts, axis=1) if parts else np.zeros((len(entries), 0), np.int8)), fp_ranges
    rows = b.reshape(n, plan.row_size)[entries] if n else np.zeros((0, plan.row_size), np.int8)
    for s in fp_sum:
        fp_ranges.append((plan.slot_offset[s], plan.slot_offset[s] + 8))
    return rows, fp_ranges


def compact_buffer_equal(got, gplan, want, wplan, perm, fp_rtol=FP_RTOL):
    """Entry i of the compacted/sorted product buffer must equal entry perm[i] of the oracle's full buffer."""
    g, fr = _entry_bytes(got, gplan, np.arange(gplan.entry_count))
    w, _ = _entry_bytes(want, wplan, perm)
    assert g.shape == w.shape, (g.shape, w.shape)
    mask = np.ones(g.shape[1], dtype=bool)
    for lo, hi in fr:
        mask[lo:hi] = False
        a = np.ascontiguousarray(g[:, lo:hi]).view(np.float64).ravel()
        b = np.ascontiguousarray(w[:, lo:hi]).view(np.float64).ravel()
        ok = (a == b) | (np.abs(a - b) <= fp_rtol * np.abs(b))
        assert ok.all(), f"fp SUM bytes {lo}:{hi}: {a[~ok][:4]} vs {b[~ok][:4]}"
    assert np.array_equal(g[:, mask], w[:, mask]), "sorted/compacted buffer differs from the oracle's entries"


def buffers_equal(got: np.ndarray, want: np.ndarray, plan: abi.Plan, fp_rtol=FP_RTOL, empty=None, float_atol=0.0):
    """Raw output buffers in the reference's row-wise layout.  Integer/bit-pattern slots must be identical; slots that
    hold a floating-point SUM (order of additions differs on a GPU) are compared within fp_rtol."""
    assert got.size == want.size == plan.buffer_size
    if plan.buffer_size == 0:
        return
    if plan.output_columnar:
        return columnar_buffers_equal(got, want, plan, fp_rtol, empty, float_atol)
    rs = plan.row_size
    g = got.view(np.int8).reshape(-1, rs)
    w = want.view(np.int8).reshape(-1, rs)
    fp_sum_slots = set()
    for t in plan.targets[: plan.num_targets]:
        if t.is_agg and t.agg_kind in (abi.kSUM, abi.kAVG) and t.agg_arg_type.type == abi.kDOUBLE:
            fp_sum_slots.add(t.first_slot)
    mask = np.ones(rs, dtype=bool)
    for t in plan.targets[: plan.num_targets]:   # SUM / AVG of a FLOAT: a float32 in the slot's low 4 bytes, compared loosely
        if t.is_agg and t.agg_kind in (abi.kSUM, abi.kAVG) and t.agg_arg_type.type == abi.kFLOAT:
            off = plan.slot_offset[t.first_slot]
            mask[off:off + 4] = False
            a = np.ascontiguousarray(g[:, off:off + 4]).view(np.float32).ravel().astype(np.float64)
            b = np.ascontiguousarray(w[:, off:off + 4]).view(np.float32).ravel().astype(np.float64)
            if empty is not None:
                a, b = a[~empty], b[~empty]
            ok = (a == b) | (np.abs(a - b) <= FLOAT_SUM_RTOL * np.abs(b) + float_atol)
            assert ok.all(), f"float SUM slot {t.first_slot}: {a[~ok][:4]} vs {b[~ok][:4]}"
    for s in fp_sum_slots:
        off = plan.slot_offset[s]
        mask[off:off + 8] = False
        a = np.ascontiguousarray(g[:, off:off + 8]).view(np.float64).ravel()
        b = np.ascontiguousarray(w[:, off:off + 8]).view(np.float64).ravel()
        if empty is not None:
            a, b = a[~empty], b[~empty]
        ok = (a == b) | (np.abs(a - b) <= fp_rtol * np.abs(b))
        assert ok.all(), f"fp SUM slot {s}: {a[~ok][:4]} vs {b[~ok][:4]}"
    if empty is not None:
        # entries every reader treats as empty: only the emptiness itself must agree (checked by the caller)
        g, w = g[~empty], w[~empty]
    assert np.array_equal(g[:, mask], w[:, mask]), "integer part of the output buffer differs from the oracle"


def columnar_buffers_equal(got, want, plan, fp_rtol, empty, float_atol=0.0):
    """Columnar layout (ResultSet.h:72-84): int64 key columns (unless keyless), then one 8-byte-aligned column per slot."""
    n = plan.entry_count
    g, w = got.view(np.int8), want.view(np.int8)
    keep = slice(None) if empty is None else ~empty
    if plan.query_desc_type != abi.NonGroupedAggregate and not plan.keyless_hash:
        stride = (8 * n + 7) // 8 * 8
        for c in range(max(plan.num_group_cols, 1)):
            a = g[c * stride:c * stride + 8 * n].view(np.int64)
            b = w[c * stride:c * stride + 8 * n].view(np.int64)
            assert np.array_equal(a[keep], b[keep]), f"key column {c} differs from the oracle"
    fp_sum_slots = {t.first_slot for t in plan.targets[: plan.num_targets]
                    if t.is_agg and t.agg_kind in (abi.kSUM, abi.kAVG) and t.agg_arg_type.type == abi.kDOUBLE}
    float_sum_slots = {t.first_slot for t in plan.targets[: plan.num_targets]
                       if t.is_agg and t.agg_kind in (abi.kSUM, abi.kAVG) and t.agg_arg_type.type == abi.kFLOAT}
    for s in range(plan.num_slots):
        wd = plan.slot_padded_width[s]
        if wd == 0:
            continue
        off = plan.slot_offset[s]
        dt = np.int32 if wd == 4 else np.int64
        a, b = g[off:off + wd * n].view(dt)[keep], w[off:off + wd * n].view(dt)[keep]
        if s in fp_sum_slots:
            a, b = a.view(np.float64), b.view(np.float64)
            ok = (a == b) | (np.abs(a - b) <= fp_rtol * np.abs(b))
            assert ok.all(), f"fp SUM slot {s}: {a[~ok][:4]} vs {b[~ok][:4]}"
        elif s in float_sum_slots:
            a32 = np.ascontiguousarray(a).view(np.float32)[0::2].astype(np.float64)
            b32 = np.ascontiguousarray(b).view(np.float32)[0::2].astype(np.float64)
            ok = (a32 == b32) | (np.abs(a32 - b32) <= FLOAT_SUM_RTOL * np.abs(b32) + float_atol)
            assert ok.all(), f"float SUM slot {s}: {a32[~ok][:4]} vs {b32[~ok][:4]}"
            assert np.array_equal(np.ascontiguousarray(a).view(np.int32)[1::2], np.ascontiguousarray(b).view(np.int32)[1::2])
        else:
            assert np.array_equal(a, b), f"slot column {s} differs from the oracle"
    if empty is None:
        mask = np.ones(g.size, dtype=bool)      # whole buffer incl. alignment padding, minus the fp SUM columns
        for s in fp_sum_slots | float_sum_slots:
            mask[plan.slot_offset[s]:plan.slot_offset[s] + 8 * n] = False
        assert np.array_equal(g[mask], w[mask]), "columnar buffer differs from the oracle outside the fp SUM columns"


def run_both(unit: abi.BuiltUnit, table: abi.Table, entry_guess=0, has_card=False, bigint_count=False, force_kernel=0,
             device_resident=True, compare_buffers=True, oracle_threads=4, dev_table: DeviceTable | None = None,
             output_columnar=False):
    """Returns (gpu ResultSet, oracle result).  Asserts plan, rows and (unless baseline) buffers agree."""
    ex = executor.Executor()
    eo = executor.execution_options(bigint_count=bigint_count, force_kernel=force_kernel, output_columnar_hint=output_columnar)
    if device_resident:
        dt = dev_table or DeviceTable(table)
        rs = ex.executeWorkUnit(entry_guess, True, dt.table, unit, eo=eo, has_cardinality_estimation=has_card,
                                memory_level=abi.GPU_LEVEL)
    else:
        rs = ex.executeWorkUnit(entry_guess, True, table, unit, eo=eo, has_cardinality_estimation=has_card,
                                memory_level=abi.CPU_LEVEL)
    ref = oracle_lib.execute(unit, table, entry_guess=entry_guess, has_card=has_card, bigint_count=bigint_count,
                             num_threads=oracle_threads, output_columnar=output_columnar)
    gp, op = rs.getQueryMemDesc(), ref.plan
    assert gp.as_dict() == op.as_dict()
    assert rs.colCount() == ref.col_count()
    for i in range(rs.colCount()):
        assert rs.getColType(i) == ref.col_type(i)
    n_rows = sum(f.num_tuples for f in table.fragments)
    rows_equal(rs.rows(), ref.rows(), col_tol=column_tolerances(gp, n_rows))
    assert rs.rowCount() == ref.row_count()
    if compare_buffers and gp.query_desc_type != abi.GroupByBaselineHash:
        n = rs.entryCount()
        assert n == ref.entry_count()
        L = oracle_lib.lib()
        if n <= 200_000:
            empty = np.array([rs.isRowAtEmpty(i) for i in range(n)], dtype=bool)
            empty_ref = np.array([bool(L.oracle_result_is_row_at_empty(ref.h, i)) for i in range(n)], dtype=bool)
            assert np.array_equal(empty, empty_ref)
        else:
            empty = None
        buffers_equal(rs.getStorageBuffer(), ref.buffer(), gp, empty=empty, float_atol=ref_tables.float_sum_atol(n_rows))
    return rs, ref
