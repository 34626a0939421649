"""Pins the oracle's hash + probe against the reference's OWN runtime compiled verbatim from /root/reference
(oracle/_ref/libref_groupby.so = QueryEngine/MurmurHash.cpp + QueryEngine/GroupByRuntime.cpp, see
oracle/ref_shim.cpp), and against the probe constants recorded in SURVEY.md §8c."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_groupby.so")


def test_murmur3_known_answers():
    L = oracle_lib.lib()
    k64 = np.array([12345], dtype=np.int64)
    k32 = np.array([7], dtype=np.int32)
    assert L.oracle_murmur3(k64.ctypes.data, 8, 0) == 342635441
    assert L.oracle_murmur3(k32.ctypes.data, 4, 0) == 1343918321


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    R = C.CDLL(REF_SO)
    R.MurmurHash3.restype = C.c_uint32
    R.MurmurHash3.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    R.get_group_value.restype = C.c_void_p
    R.get_group_value.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    R.get_group_value_fast.restype = C.c_void_p
    R.get_group_value_fast.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint32]
    return R


def test_murmur3_matches_reference(ref):
    L = oracle_lib.lib()
    rng = np.random.default_rng(7)
    data = rng.integers(0, 256, size=64, dtype=np.uint8)
    for length in range(0, 33):
        for seed in (0, 1, 0x9747B28C):
            assert L.oracle_murmur3(data.ctypes.data, length, seed) == ref.MurmurHash3(data.ctypes.data, length, seed)
    keys = rng.integers(-2**62, 2**62, size=2000, dtype=np.int64)
    for i in range(keys.size):
        p = keys[i:i + 1].ctypes.data
        assert L.oracle_murmur3(p, 8, 0) == ref.MurmurHash3(p, 8, 0)


@pytest.mark.parametrize("key_width,entry_count,nkeys", [(8, 97, 60), (8, 64, 64), (4, 101, 80), (8, 16, 40)])
def test_get_group_value_matches_reference(ref, key_width, entry_count, nkeys):
    """Same insert sequence into two tables -> identical slot addresses, identical key placement, and the same
    'table full' answer (NULL) once entry_count distinct keys are in (GroupByRuntime.cpp:25-48)."""
    L = oracle_lib.lib()
    rng = np.random.default_rng(entry_count)
    row_size_quad = 3  # [key 8][slot][slot]
    empty = np.iinfo(np.int64).max if key_width == 8 else None
    def fresh():
        b = np.zeros(entry_count * row_size_quad, dtype=np.int64)
        rows = b.reshape(entry_count, row_size_quad)
        if key_width == 8:
            rows[:, 0] = empty
        else:
            b.view(np.int32).reshape(entry_count, 2 * row_size_quad)[:, 0] = np.iinfo(np.int32).max
        return b
    ours, theirs = fresh(), fresh()
    keys = rng.integers(1, 50 if entry_count == 16 else 10**6, size=nkeys)
    keys = np.concatenate([keys, keys[: nkeys // 2]])  # revisit
    for k in keys:
        kb = np.zeros(1, dtype=np.int64)
        if key_width == 8:
            kb[0] = k
        else:
            kb.view(np.int32)[0] = k
        o = L.oracle_get_group_value(ours.ctypes.data, entry_count, kb.ctypes.data, 1, key_width, row_size_quad)
        r = ref.get_group_value(theirs.ctypes.data, entry_count, kb.ctypes.data, 1, key_width, row_size_quad)
        r_off = -1 if not r else (r - theirs.ctypes.data) // 8
        assert o == r_off
        if o >= 0:
            ours[o] += 1
            theirs[r_off] += 1
    assert np.array_equal(ours, theirs)


def test_get_group_value_fast_reference(ref):
    """Perfect-hash direct index: off = (key - min) / bucket * row_size_quad, key written on first touch
    (GroupByRuntime.cpp:194-209) — the formula the oracle's run_fragment and the CUDA kernels use."""
    row_size_quad, n = 3, 20
    buf = np.full(n * row_size_quad, np.iinfo(np.int64).max, dtype=np.int64)
    for key, mn, bucket in [(5, 0, 0), (17, 3, 0), (40, 10, 2)]:
        r = ref.get_group_value_fast(buf.ctypes.data, key, mn, bucket, row_size_quad)
        d = key - mn
        if bucket:
            d //= bucket
        assert (r - buf.ctypes.data) // 8 == d * row_size_quad + 1
        assert buf[d * row_size_quad] == key
