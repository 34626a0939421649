"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from heavydb_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_set_filter_on_deleted_column.argtypes = [C.c_int32]
        L.oracle_murmur3.restype = C.c_uint32
        L.oracle_murmur3.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.oracle_get_group_value.restype = C.c_int64
        L.oracle_get_group_value.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.oracle_plan.restype = C.c_int32
        L.oracle_plan.argtypes = [C.POINTER(abi.ExecUnit), C.POINTER(abi.TableInfo), C.POINTER(abi.ExecutionOptions),
                                  C.c_size_t, C.c_int32, C.POINTER(abi.Plan)]
        L.oracle_execute.restype = C.c_int32
        L.oracle_execute.argtypes = [C.POINTER(abi.ExecUnit), C.POINTER(abi.TableInfo), C.POINTER(abi.ExecutionOptions),
                                     C.c_size_t, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.oracle_result_plan.restype = C.POINTER(abi.Plan)
        L.oracle_result_plan.argtypes = [C.c_void_p]
        L.oracle_result_buffer.restype = C.c_void_p
        L.oracle_result_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        for name in ("oracle_result_entry_count", "oracle_result_row_count", "oracle_result_col_count"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_result_is_row_at_empty.restype = C.c_int32
        L.oracle_result_is_row_at_empty.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_result_move_to_begin.argtypes = [C.c_void_p]
        L.oracle_result_free.argtypes = [C.c_void_p]
        L.oracle_result_col_type.restype = abi.TypeInfo
        L.oracle_result_col_type.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_result_get_next_row.restype = C.c_int32
        L.oracle_result_get_next_row.argtypes = [C.c_void_p, C.POINTER(abi.TargetValue), C.c_int32]
        L.oracle_result_ndv_estimator.restype = C.c_size_t
        L.oracle_result_ndv_estimator.argtypes = [C.c_void_p]
        L.oracle_result_sort.restype = C.c_int32
        L.oracle_result_sort.argtypes = [C.c_void_p, C.POINTER(abi.OrderEntry), C.c_int32, C.c_size_t]
        L.oracle_result_drop_first_n.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_result_keep_first_n.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_result_permutation_at.restype = C.c_int64
        L.oracle_result_permutation_at.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_gen_column.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int32]
        L.oracle_gen_column_strided.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_int64,
                                                C.c_int64, C.c_int64, C.c_int32]
        L.oracle_execute_generated.restype = C.c_int32
        L.oracle_execute_generated.argtypes = [C.POINTER(abi.ExecUnit), C.POINTER(abi.TableInfo), C.POINTER(abi.ExecutionOptions),
                                               C.c_size_t, C.c_int32, C.c_int32, C.c_uint64, C.POINTER(GenCol), C.c_int64,
                                               C.POINTER(C.c_void_p)]
        L.oracle_set_thread_pinning.argtypes = [C.c_int32]
        L.oracle_gen_fragments.argtypes = [C.POINTER(C.c_void_p), C.POINTER(GenCol), C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                           C.c_int32, C.c_uint64, C.c_int32]
        L.oracle_result_from_storage.restype = C.c_int32
        L.oracle_result_from_storage.argtypes = [C.POINTER(abi.ExecUnit), C.POINTER(abi.TableInfo), C.POINTER(abi.ExecutionOptions),
                                                 C.c_size_t, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.oracle_result_sets_reduce.restype = C.c_int32
        L.oracle_result_sets_reduce.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


class GenCol(C.Structure):
    """One column of a generated table (oracle.cpp OracleGenCol): the PHYSICAL element type and the generator's parameters."""
    _fields_ = [("sql_type", C.c_int32), ("col_tag", C.c_uint32), ("lo", C.c_int64), ("span", C.c_int64), ("stride", C.c_int64)]


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


class OracleResult:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_result_free(self.h)
            self.h = None

    @property
    def plan(self) -> abi.Plan:
        return lib().oracle_result_plan(self.h).contents

    def row_count(self):
        return lib().oracle_result_row_count(self.h)

    def ndv_estimator(self):
        return lib().oracle_result_ndv_estimator(self.h)

    def sort(self, order_entries, top_n=0):
        """ResultSet::sort: order_entries = [(tle_no, is_desc, nulls_first)]."""
        arr = (abi.OrderEntry * max(len(order_entries), 1))()
        for i, (tle, desc, nf) in enumerate(order_entries):
            arr[i].tle_no, arr[i].is_desc, arr[i].nulls_first = tle, int(desc), int(nf)
        rc = lib().oracle_result_sort(self.h, arr, len(order_entries), top_n)
        if rc:
            raise OracleError(rc, "bad order entry")

    def drop_first_n(self, n):
        lib().oracle_result_drop_first_n(self.h, n)

    def keep_first_n(self, n):
        lib().oracle_result_keep_first_n(self.h, n)

    def permutation(self):
        return [lib().oracle_result_permutation_at(self.h, i) for i in range(self.entry_count())]

    def col_count(self):
        return lib().oracle_result_col_count(self.h)

    def entry_count(self):
        return lib().oracle_result_entry_count(self.h)

    def col_type(self, i):
        t = lib().oracle_result_col_type(self.h, i)
        return (t.type, t.notnull, t.scale)

    def buffer(self) -> np.ndarray:
        n = C.c_size_t()
        p = lib().oracle_result_buffer(self.h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=np.int8)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=(n.value,)).copy()

    def rows(self, decimal_to_double=True):
        """All rows via getNextRow, as tuples of python values (None = NULL)."""
        L = lib()
        L.oracle_result_move_to_begin(self.h)
        nc = self.col_count()
        row = (abi.TargetValue * nc)()
        out = []
        while L.oracle_result_get_next_row(self.h, row, int(decimal_to_double)):
            out.append(tuple(v.py() for v in row))
        return out


def make_eo(bigint_count=False, force_kernel=0, output_columnar=False, device=-1):
    eo = abi.ExecutionOptions()
    eo.allow_multifrag = 1
    eo.output_columnar_hint = int(output_columnar)
    eo.bigint_count = int(bigint_count)
    eo.force_kernel = force_kernel
    eo.device_ordinal = device
    return eo


def plan(unit: abi.BuiltUnit, table: abi.Table, entry_guess=0, has_card=False, bigint_count=False,
         output_columnar=False) -> abi.Plan:
    bt = table.build(abi.CPU_LEVEL)
    out = abi.Plan()
    eo = make_eo(bigint_count, output_columnar=output_columnar)
    rc = lib().oracle_plan(C.byref(unit.unit), C.byref(bt.info), C.byref(eo), entry_guess, int(has_card), C.byref(out))
    if rc:
        raise OracleError(rc, lib().oracle_last_error().decode())
    return out


def execute(unit: abi.BuiltUnit, table: abi.Table, entry_guess=0, has_card=False, bigint_count=False,
            num_threads=1, output_columnar=False) -> OracleResult:
    bt = table.build(abi.CPU_LEVEL)
    h = C.c_void_p()
    eo = make_eo(bigint_count, output_columnar=output_columnar)
    rc = lib().oracle_execute(C.byref(unit.unit), C.byref(bt.info), C.byref(eo), entry_guess, int(has_card),
                              num_threads, C.byref(h))
    if rc:
        raise OracleError(rc, lib().oracle_last_error().decode())
    return OracleResult(h)


def execute_generated(unit: abi.BuiltUnit, table: abi.Table, gen_cols, seed, rows_per_fragment_id, entry_guess=0, has_card=False,
                      bigint_count=False, num_threads=1, output_columnar=False) -> OracleResult:
    """The oracle over a table that is generated on the fly (never materialised): `table` carries fragment sizes, ids and
    chunk stats only; gen_cols = [(physical sql type, col_tag, lo, span, stride)] per column; global row of tuple i of
    fragment f = fragment_id * rows_per_fragment_id + i.  One output buffer per worker thread (multi-fragment kernels)."""
    bt = table.build(abi.CPU_LEVEL)
    arr = (GenCol * len(gen_cols))(*[GenCol(int(t), int(tag), int(lo), int(span), int(stride)) for t, tag, lo, span, stride in gen_cols])
    h = C.c_void_p()
    eo = make_eo(bigint_count, output_columnar=output_columnar)
    rc = lib().oracle_execute_generated(C.byref(unit.unit), C.byref(bt.info), C.byref(eo), entry_guess, int(has_card), num_threads,
                                        seed, arr, rows_per_fragment_id, C.byref(h))
    if rc:
        raise OracleError(rc, lib().oracle_last_error().decode())
    return OracleResult(h)


def result_from_storage(unit: abi.BuiltUnit, table: abi.Table, storage: np.ndarray, entry_guess=0, has_card=False,
                        bigint_count=False, output_columnar=False) -> OracleResult:
    """ResultSet(targets, device_type, query_mem_desc, ...) + allocateStorage() over bytes the caller filled
    (Tests/ResultSetTest.cpp:879-894): the descriptor is the planned query's."""
    bt = table.build(abi.CPU_LEVEL)
    h = C.c_void_p()
    eo = make_eo(bigint_count, output_columnar=output_columnar)
    buf = np.ascontiguousarray(storage).view(np.uint8)
    rc = lib().oracle_result_from_storage(C.byref(unit.unit), C.byref(bt.info), C.byref(eo), entry_guess, int(has_card),
                                          buf.ctypes.data, buf.size, C.byref(h))
    if rc:
        raise OracleError(rc, lib().oracle_last_error().decode())
    return OracleResult(h)


def reduce_result_sets(results) -> OracleResult:
    """ResultSetManager::reduce (ResultSetReduction.cpp:1055-1140) over result sets of one descriptor."""
    arr = (C.c_void_p * len(results))(*[r.h for r in results])
    h = C.c_void_p()
    rc = lib().oracle_result_sets_reduce(arr, len(results), C.byref(h))
    if rc:
        raise OracleError(rc, lib().oracle_last_error().decode())
    return OracleResult(h)


def set_thread_pinning(on: bool):
    """Worker t of execute() / gen_fragments() runs on the t-th CPU of the process's affinity mask (NUMA-local first touch)."""
    lib().oracle_set_thread_pinning(int(on))


def gen_fragments(frag_arrays, gen_cols, rows, row0, seed, num_threads):
    """Fill whole fragments (frag_arrays[f][c]: preallocated, untouched numpy arrays) with the counter-based generator, fragment f
    by worker f % num_threads — the same worker execute() hands fragment f to."""
    nf, nc = len(frag_arrays), len(gen_cols)
    ptrs = (C.c_void_p * (nf * nc))(*[a.ctypes.data for fr in frag_arrays for a in fr])
    arr = (GenCol * nc)(*[GenCol(int(t), int(tag), int(lo), int(span), int(stride)) for t, tag, lo, span, stride in gen_cols])
    lib().oracle_gen_fragments(ptrs, arr, nc, (C.c_int64 * nf)(*rows), (C.c_int64 * nf)(*row0), nf, seed, num_threads)


def gen_column(sql_type, seed, col_tag, row0, count, lo=0, span=1, threads=8, stride=1) -> np.ndarray:
    a = np.empty(count, dtype=abi.NUMPY_OF[sql_type])
    lib().oracle_gen_column_strided(a.ctypes.data, sql_type, seed, col_tag, row0, count, lo, span, stride, threads)
    return a
