/*
 * kernels.cu — hand-written sm_100a kernels for the scan -> filter -> group-by/aggregate path.
 *
 * They replace what the reference JIT-compiles per query:
 *   row loop            query_group_by_template / query_template + multifrag_query_hoisted_literals
 *                       (QueryEngine/QueryTemplateGenerator.cpp:552-815,257-549; RuntimeFunctions.cpp:2434-2472)
 *   column decode       fixed_width_int_decode / fixed_width_double_decode (QueryEngine/DecodersImpl.h:30-61,112-136)
 *   filter              DEF_CMP_NULLABLE + toBool (RuntimeFunctions.cpp:73-107, LogicalIR.cpp:344-352)
 *   group lookup        get_group_value_fast[_keyless] (GroupByRuntime.cpp:194-209, RuntimeFunctions.cpp:2126-2152),
 *                       get_group_value + get_matching_group_value (GroupByRuntime.cpp:20-48, cuda_mapd_rt.cu:180-216)
 *   aggregate update    agg_*_shared / agg_*_skip_val_shared (cuda_mapd_rt.cu:437-1198)
 *   smem table          init_shared_mem + JIT'd reduce_from_smem_to_gmem (cuda_mapd_rt.cu:73-87,
 *                       GpuSharedMemoryUtils.cpp:96-383)
 *   buffer init         init_group_by_buffer_gpu (GpuInitGroups.cu:124-171)
 *
 * Design (DESIGN.md has the numbers):
 *   - persistent CTAs, grid = #SMs x CTAs/SM, static chunk striding over all fragments of the launch;
 *   - each thread owns R rows per chunk, lane-consecutive => every column load is a fully coalesced
 *     ld.global.nc.L1::no_allocate of the column's own width; R independent loads per column are in flight;
 *   - vector-at-a-time interpretation of the device program: all operator/width switches are warp-uniform and
 *     executed once per R rows;
 *   - group table private to the CTA in shared memory (TMA bulk copy of an identity image initialises it),
 *     warp-private replicas for small tables; 32-bit native shared atomics only: a 64-bit integer SUM keeps its low
 *     word in shared memory and sends the (rare) carries straight to the HBM table, because sm_100a has no native
 *     64-bit shared-memory atomic add (ATOMS.CAST.SPIN loops otherwise);
 *   - warp-aggregated update (shuffle reduction, one atomic per warp) for the non-grouped case;
 *   - tables too large for shared memory go to one dense table in HBM/L2 with RED.E.ADD/MIN/MAX;
 *   - sparse keys: open addressing in HBM, MurmurHash3 (same function and home slot as the reference), 64-bit CAS.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "b2q_internal.h"

namespace b2q {

constexpr int R = 8;                 /* rows per thread per chunk */
constexpr int kMaxBlock = 1024;

enum { MODE_SMEM = 0, MODE_GLOBAL = 1, MODE_BASELINE = 2 };

/* ---------------------------------------------------------------------------------------------------------
 * loads: streaming (read-only path, no L1 allocation), branch-free predication.
 * The asm is deliberately NOT volatile: the data is immutable for the kernel, so the compiler may hoist and batch
 * the loads of a vector, which is what puts R independent requests per column in flight.
 * ------------------------------------------------------------------------------------------------------- */
template <bool PRED>
__device__ __forceinline__ int64_t ldg_b64(const int8_t* p, uint32_t pred, uint64_t pol) {
  int64_t v;
  if (PRED)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b64 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %3;\n\t}" : "=l"(v) : "l"(p), "r"(pred), "l"(pol));
  else
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
template <bool PRED>
__device__ __forceinline__ int32_t ldg_s32(const int8_t* p, uint32_t pred, uint64_t pol) {
  int32_t v;
  if (PRED)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %3;\n\t}" : "=r"(v) : "l"(p), "r"(pred), "l"(pol));
  else
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
template <bool PRED>
__device__ __forceinline__ int32_t ldg_s16(const int8_t* p, uint32_t pred, uint64_t pol) {
  int32_t v;
  if (PRED)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.s16 %0, [%1], %3;\n\t}" : "=r"(v) : "l"(p), "r"(pred), "l"(pol));
  else
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s16 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return (int32_t)(int16_t)v;
}
template <bool PRED>
__device__ __forceinline__ int32_t ldg_s8(const int8_t* p, uint32_t pred, uint64_t pol) {
  int32_t v;
  if (PRED)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.s8 %0, [%1], %3;\n\t}" : "=r"(v) : "l"(p), "r"(pred), "l"(pol));
  else
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return (int32_t)(int8_t)v;
}

/* zero-extending variants: dictionary ids stored on 1 / 2 bytes are unsigned (FixedWidthUnsigned, ColumnIR.cpp:59-67) */
template <bool PRED>
__device__ __forceinline__ int32_t ldg_u16(const int8_t* p, uint32_t pred, uint64_t pol) {
  uint32_t v;
  if (PRED)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %3;\n\t}" : "=r"(v) : "l"(p), "r"(pred), "l"(pol));
  else
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return (int32_t)(v & 0xFFFFu);
}
template <bool PRED>
__device__ __forceinline__ int32_t ldg_u8(const int8_t* p, uint32_t pred, uint64_t pol) {
  uint32_t v;
  if (PRED)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b32 %0, 0;\n\t@p ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %3;\n\t}" : "=r"(v) : "l"(p), "r"(pred), "l"(pol));
  else
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return (int32_t)(v & 0xFFu);
}

/* R rows of an 8-byte column: rows row0 + j*stride */
/* jidx != nullptr: the column belongs to the joined inner table and is read at the matching inner rows (a gather
 * through the normal cached path: dimension tables are small and re-read constantly).  Call sites pass a
 * compile-time nullptr in the kernels without a join level, so the branch disappears there. */
template <bool PRED>
__device__ __forceinline__ void load64(int64_t (&v)[R], const int8_t* __restrict__ base, int64_t row0, int stride, uint32_t mask, uint64_t pol,
                                       const int32_t* jidx = nullptr, const int32_t* /* jval: only 1/2/4-byte columns are packed */ = nullptr,
                                       const int64_t* nullp = nullptr) {
  if (jidx) { /* nullp (LEFT join kernels only): an unmatched row (idx < 0) reads NULL (codegenOuterJoinNullPlaceholder) */
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = (mask >> j & 1) ? ((nullp && jidx[j] < 0) ? *nullp : __ldg(reinterpret_cast<const long long*>(base) + jidx[j])) : 0;
    return;
  }
  const int8_t* p = base + row0 * 8;
  const int64_t step = (int64_t)stride * 8;
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = ldg_b64<PRED>(p + j * step, mask >> j & 1, pol);
}
/* R rows of a 1/2/4-byte integer column, sign-extended to 32 bits (width -1 / -2: zero-extended) */
template <bool PRED>
__device__ __forceinline__ void load32(int32_t (&v)[R], const int8_t* __restrict__ base, int width, int64_t row0, int stride, uint32_t mask, uint64_t pol,
                                       const int32_t* jidx = nullptr, const int32_t* jval = nullptr, const int64_t* nullp = nullptr) {
  if (jval) { /* the column that rides in the packed join table: already in registers since the probe */
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = (mask >> j & 1) ? jval[j] : 0;
    return;
  }
  if (jidx) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int32_t x = 0;
      if (nullp && (mask >> j & 1) && jidx[j] < 0) x = (int32_t)*nullp;
      else if (mask >> j & 1) {
        const int64_t i = jidx[j];
        switch (width) {
          case 4: x = __ldg(reinterpret_cast<const int32_t*>(base) + i); break;
          case 2: x = __ldg(reinterpret_cast<const int16_t*>(base) + i); break;
          case -2: x = __ldg(reinterpret_cast<const uint16_t*>(base) + i); break;
          case -1: x = __ldg(reinterpret_cast<const uint8_t*>(base) + i); break;
          default: x = __ldg(reinterpret_cast<const signed char*>(base) + i); break;
        }
      }
      v[j] = x;
    }
    return;
  }
  if (width == 4) {
    const int8_t* p = base + row0 * 4;
    const int64_t step = (int64_t)stride * 4;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = ldg_s32<PRED>(p + j * step, mask >> j & 1, pol);
  } else if (width == 2) {
    const int8_t* p = base + row0 * 2;
    const int64_t step = (int64_t)stride * 2;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = ldg_s16<PRED>(p + j * step, mask >> j & 1, pol);
  } else if (width == -2) {
    const int8_t* p = base + row0 * 2;
    const int64_t step = (int64_t)stride * 2;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = ldg_u16<PRED>(p + j * step, mask >> j & 1, pol);
  } else if (width == -1) {
    const int8_t* p = base + row0;
    const int64_t step = (int64_t)stride;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = ldg_u8<PRED>(p + j * step, mask >> j & 1, pol);
  } else {
    const int8_t* p = base + row0;
    const int64_t step = (int64_t)stride;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = ldg_s8<PRED>(p + j * step, mask >> j & 1, pol);
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * filter: one comparison = one unsigned range test, (v - lo) <= span, in the column's register class
 * ------------------------------------------------------------------------------------------------------- */
template <bool FULL>
__device__ __forceinline__ uint32_t eval_term(const DevTerm& t, const int8_t* const* __restrict__ cols, int64_t row0,
                                              int stride, uint32_t valid, uint64_t pol, const int32_t* jidx = nullptr,
                                              const int32_t* jval = nullptr, const int64_t* jnull = nullptr) {
  uint32_t m = 0;
  const bool neg = t.negate;
  if (!t.cmp_fp) {
    if (t.width == 8) {
      int64_t v[R];
      load64<!FULL>(v, cols[t.col], row0, stride, valid, pol, jidx, nullptr, jnull);
      const uint64_t lo = (uint64_t)t.lo, span = t.span;
      if (lo == 0x8000000000000000ull) { /* only an upper bound (`<`, `<=`): one signed compare instead of subtract + compare */
        const int64_t hi = (int64_t)(lo + span);
#pragma unroll
        for (int j = 0; j < R; ++j) m |= (uint32_t)((v[j] <= hi) != neg) << j;
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) m |= (uint32_t)(((uint64_t)v[j] - lo <= span) != neg) << j;
      }
      if (t.null_check) {
        const int64_t nullv = t.null_bits;
#pragma unroll
        for (int j = 0; j < R; ++j) m &= ~((uint32_t)(v[j] == nullv) << j);
      }
    } else {
      int32_t v[R];
      load32<!FULL>(v, cols[t.col], t.width, row0, stride, valid, pol, jidx, jval, jnull);
      const uint32_t lo = (uint32_t)t.lo, span = (uint32_t)t.span;
#pragma unroll
      for (int j = 0; j < R; ++j) m |= (uint32_t)(((uint32_t)v[j] - lo <= span) != neg) << j;
      if (t.null_check) {
        const int32_t nullv = (int32_t)t.null_bits;
#pragma unroll
        for (int j = 0; j < R; ++j) m &= ~((uint32_t)(v[j] == nullv) << j);
      }
    }
  } else {
    const double lo = t.flo, hi = t.fhi;
    double d[R];
    uint32_t isnull = 0;
    if (t.col_is_fp) {
      int64_t v[R];
      load64<!FULL>(v, cols[t.col], row0, stride, valid, pol, jidx, nullptr, jnull);
      const double nullv = __longlong_as_double(t.null_bits);
#pragma unroll
      for (int j = 0; j < R; ++j) { d[j] = __longlong_as_double(v[j]); isnull |= (uint32_t)(d[j] == nullv) << j; }
    } else if (t.width == 8) {
      int64_t v[R];
      load64<!FULL>(v, cols[t.col], row0, stride, valid, pol, jidx, nullptr, jnull);
      const int64_t nullv = t.null_bits;
#pragma unroll
      for (int j = 0; j < R; ++j) { d[j] = (double)v[j]; isnull |= (uint32_t)(v[j] == nullv) << j; }
    } else {
      int32_t v[R];
      load32<!FULL>(v, cols[t.col], t.width, row0, stride, valid, pol, jidx, jval, jnull);
      const int32_t nullv = (int32_t)t.null_bits;
#pragma unroll
      for (int j = 0; j < R; ++j) { d[j] = (double)v[j]; isnull |= (uint32_t)(v[j] == nullv) << j; }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) m |= (uint32_t)(((d[j] >= lo) & (d[j] <= hi)) != neg) << j;
    if (t.null_check) m &= ~isnull;
  }
  return m & valid;
}

/* column OP column */
template <bool FULL>
__device__ __forceinline__ uint32_t eval_term2(const DevTerm& t, const int8_t* const* __restrict__ cols, int64_t row0, int stride,
                                               uint32_t valid, uint64_t pol, const int32_t* jidx1, const int32_t* jval1,
                                               const int64_t* jnull1, const int32_t* jidx2, const int32_t* jval2, const int64_t* jnull2) {
  int64_t a[R], b[R];
  if (t.width == 8) load64<!FULL>(a, cols[t.col], row0, stride, valid, pol, jidx1, jval1, jnull1);
  else {
    int32_t x[R];
    load32<!FULL>(x, cols[t.col], t.width, row0, stride, valid, pol, jidx1, jval1, jnull1);
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = x[j];
  }
  if (t.width2 == 8) load64<!FULL>(b, cols[t.col2], row0, stride, valid, pol, jidx2, jval2, jnull2);
  else {
    int32_t x[R];
    load32<!FULL>(x, cols[t.col2], t.width2, row0, stride, valid, pol, jidx2, jval2, jnull2);
#pragma unroll
    for (int j = 0; j < R; ++j) b[j] = x[j];
  }
  uint32_t isnull = 0, m = 0;
  const int op = t.op2;
  if (t.cmp_fp) {
    const double n1 = __longlong_as_double(t.null_bits), n2 = __longlong_as_double(t.null_bits2);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      double x, y;
      if (t.col_is_fp) { x = __longlong_as_double(a[j]); isnull |= (uint32_t)(t.nullable1 && x == n1) << j; }
      else { x = (double)a[j]; isnull |= (uint32_t)(t.nullable1 && a[j] == t.null_bits) << j; }
      if (t.col2_is_fp) { y = __longlong_as_double(b[j]); isnull |= (uint32_t)(t.nullable2 && y == n2) << j; }
      else { y = (double)b[j]; isnull |= (uint32_t)(t.nullable2 && b[j] == t.null_bits2) << j; }
      const bool r = op == B2Q_kEQ ? x == y : op == B2Q_kNE ? x != y : op == B2Q_kLT ? x < y : op == B2Q_kGT ? x > y : op == B2Q_kLE ? x <= y : x >= y;
      m |= (uint32_t)r << j;
    }
  } else {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      isnull |= (uint32_t)((t.nullable1 && a[j] == t.null_bits) || (t.nullable2 && b[j] == t.null_bits2)) << j;
      const bool r = op == B2Q_kEQ ? a[j] == b[j] : op == B2Q_kNE ? a[j] != b[j] : op == B2Q_kLT ? a[j] < b[j] : op == B2Q_kGT ? a[j] > b[j] : op == B2Q_kLE ? a[j] <= b[j] : a[j] >= b[j];
      m |= (uint32_t)r << j;
    }
  }
  return m & ~isnull & valid;
}

template <bool FULL, int JOIN>
__device__ __forceinline__ uint32_t eval_filter(const DevFilter& f, const int8_t* const* __restrict__ cols,
                                                int64_t row0, int stride, uint32_t valid, uint64_t pol,
                                                const int8_t* __restrict__ col_inner, const int32_t* jidx, int packed_col,
                                                const int32_t* jval, const int64_t* __restrict__ col_null) {
#define B2Q_TERM_JX(t) ((JOIN && col_inner[(t).col]) ? jidx : nullptr), ((JOIN && (t).col == packed_col) ? jval : nullptr), (JOIN == 2 ? col_null + (t).col : nullptr)
#define B2Q_TERM_JX2(t) ((JOIN && col_inner[(t).col2]) ? jidx : nullptr), ((JOIN && (t).col2 == packed_col) ? jval : nullptr), (JOIN == 2 ? col_null + (t).col2 : nullptr)
#define B2Q_EVAL_TERM(t) ((t).col2 >= 0 ? eval_term2<FULL>((t), cols, row0, stride, valid, pol, B2Q_TERM_JX(t), B2Q_TERM_JX2(t)) \
                                       : eval_term<FULL>((t), cols, row0, stride, valid, pol, B2Q_TERM_JX(t)))
  if (f.n_ops == 0) return valid;
  if (f.n_ops == 1) return B2Q_EVAL_TERM(f.terms[0]);
  uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int i = 0; i < f.n_ops; ++i) {
    const uint32_t op = f.ops[i];
    const uint32_t kind = op >> 4;
    if (kind == FOP_TERM) {
      const uint32_t m = B2Q_EVAL_TERM(f.terms[op & 15]);
      s3 = s2; s2 = s1; s1 = s0; s0 = m;
    } else {
      s0 = (kind == FOP_AND) ? (s1 & s0) : (s1 | s0);
      s1 = s2; s2 = s3; s3 = 0;
    }
  }
  return s0 & valid;
#undef B2Q_EVAL_TERM
#undef B2Q_TERM_JX2
#undef B2Q_TERM_JX
}

/* ---------------------------------------------------------------------------------------------------------
 * skip test (NULL handling of aggregate arguments), see DevAcc
 * ------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t not_skipped64(const DevAcc& a, const int64_t (&v)[R], uint32_t pass) {
  if (!a.skip1_en && !a.skip2_en) return pass;
  uint32_t m = 0;
  if (a.is_fp) {
    const double s = __longlong_as_double(a.skip1_val);
#pragma unroll
    for (int j = 0; j < R; ++j) m |= (uint32_t)(__longlong_as_double(v[j]) != s) << j;
  } else {
    const int64_t s1 = a.skip1_val, s2 = a.skip2_val;
    const bool e1 = a.skip1_en, e2 = a.skip2_en, tr = a.skip2_trunc32;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t w = tr ? (int64_t)(int32_t)v[j] : v[j];
      const bool skip = (e1 & (v[j] == s1)) | (e2 & (w == s2));
      m |= (uint32_t)(!skip) << j;
    }
  }
  return m & pass;
}
__device__ __forceinline__ uint32_t not_skipped32(const DevAcc& a, const int32_t (&v)[R], uint32_t pass) {
  if (!a.skip1_en && !a.skip2_en) return pass;
  /* a sign-extended 32-bit value can only equal a skip value that itself fits in 32 bits */
  const bool e1 = a.skip1_en && a.skip1_val == (int64_t)(int32_t)a.skip1_val;
  const bool e2 = a.skip2_en && a.skip2_val == (int64_t)(int32_t)a.skip2_val;
  const int32_t s1 = (int32_t)a.skip1_val, s2 = (int32_t)a.skip2_val;
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const bool skip = (e1 & (v[j] == s1)) | (e2 & (v[j] == s2));
    m |= (uint32_t)(!skip) << j;
  }
  return m & pass;
}

/* ---------------------------------------------------------------------------------------------------------
 * MurmurHash3 x86_32 for one 4- or 8-byte key, seed 0 (QueryEngine/MurmurHash3Inl.h:11-72)
 * ------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ __forceinline__ uint32_t murmur_block(uint32_t h1, uint32_t k1) {
  k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u;
  h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
  return h1;
}
__device__ __forceinline__ uint32_t murmur3_key(int64_t key, int width) {
  uint32_t h1 = 0;
  h1 = murmur_block(h1, (uint32_t)key);
  if (width == 8) h1 = murmur_block(h1, (uint32_t)((uint64_t)key >> 32));
  h1 ^= (uint32_t)width;
  h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
  return h1;
}

/* MurmurHash3_x86_32 finalisation for a key of `len` bytes whose 4-byte blocks were folded with murmur_block */
__device__ __forceinline__ uint32_t murmur3_fmix(uint32_t h1, uint32_t len) {
  h1 ^= len;
  h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
  return h1;
}

/* ---------------------------------------------------------------------------------------------------------
 * global (HBM / L2) reductions without return value
 * ------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void red_add_u64(int64_t* p, uint64_t v) { asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_add_f64(int64_t* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void red_min_s64(int64_t* p, int64_t v) { asm volatile("red.global.min.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_max_s64(int64_t* p, int64_t v) { asm volatile("red.global.max.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

/* HBM/L2-resident table (MODE_GLOBAL / MODE_BASELINE).  COUNT and integer SUM keep a SPLIT accumulator: the array's
 * first 4n bytes are the low words, the next 4n bytes the high words.  Only the low words are hot (native 32-bit
 * L2 atomics); the high word sees the rare carry / a value wider than 32 bits.  That halves the randomly accessed
 * footprint — 40 MB instead of 80 MB for 1e7 groups — which is what lets the table stay L2-resident while 16 GB of
 * column data streams through the same L2 (ncu: with 64-bit REDs on the 80 MB table, DRAM traffic was 2.9x the
 * algorithmic bytes; see profiles/r1_scan_c4_*.txt).  Table accesses carry an evict_last L2 policy, the column
 * stream evict_first. */
__device__ __forceinline__ uint32_t global_split_add(int64_t* arr, uint32_t e, int64_t n, uint32_t vl, int32_t vh, uint64_t pol_tab) {
  uint32_t* lo = reinterpret_cast<uint32_t*>(arr) + e;
  uint32_t old;
  asm volatile("atom.global.add.L2::cache_hint.u32 %0, [%1], %2, %3;" : "=r"(old) : "l"(lo), "r"(vl), "l"(pol_tab) : "memory");
  const int32_t hi = vh + (int32_t)((uint32_t)(old + vl) < old);
  if (hi != 0) atomicAdd(reinterpret_cast<int32_t*>(arr) + n + e, hi);
  return old;
}
/* "group touched" flag piggy-backed on an accumulator that every passing row updates: the FIRST atomic on an entry
 * always returns the initial 0, so storing the flag whenever 0 comes back marks every touched group and costs no
 * extra L2 request for the (overwhelmingly common) rows that see a non-zero running value. */
__device__ __forceinline__ void global_split_add_touch(int64_t* arr, uint8_t* flags, uint32_t e, int64_t n, uint32_t vl, int32_t vh, uint64_t pol_tab) {
  const uint32_t old = global_split_add(arr, e, n, vl, vh, pol_tab);
  if (flags && old == 0) flags[e] = 1;
}

__device__ __forceinline__ void global_update(int op, int64_t* arr, uint8_t* flags, uint32_t e, int64_t n, int64_t v, uint64_t pol_tab) {
  switch (op) {
    case ACC_COUNT: global_split_add_touch(arr, flags, e, n, 1u, 0, pol_tab); break;
    case ACC_SUM_I64: global_split_add_touch(arr, flags, e, n, (uint32_t)v, (int32_t)(v >> 32), pol_tab); break;
    case ACC_SUM_F64: red_add_f64(arr + e, __longlong_as_double(v)); break;
    case ACC_MIN_I64: red_min_s64(arr + e, v); break;
    case ACC_MAX_I64: red_max_s64(arr + e, v); break;
    case ACC_MIN_F64: { const double d = __longlong_as_double(v); if (d == d) red_min_s64(arr + e, b2q_f64_to_ord(v)); break; }
    default: { const double d = __longlong_as_double(v); if (d == d) red_max_s64(arr + e, b2q_f64_to_ord(v)); break; }
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * shared-memory table updates.  `tab` points at the accumulator's array inside this warp's replica.
 * 64-bit integer SUM: (hi:lo) += v with lo in shared memory (native 32-bit ATOMS.ADD) and the rare hi deltas
 * (carry out of lo, or a value that does not fit 32 bits) sent to the HBM table with RED.ADD.64.
 * ------------------------------------------------------------------------------------------------------- */
/* branch-free predicated forms: the compiler wraps `if (p) atomicAdd(...)` in BSSY/BRA/BSYNC per row; a predicated
 * ATOMS needs none of that (profiles/r1_scan_c2_v2: 587 warp-instructions per 8-row iteration, ~70 of them for this) */
__device__ __forceinline__ void smem_inc_pred(uint32_t saddr, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(saddr), "r"(pred) : "memory");
}
__device__ __forceinline__ uint32_t smem_add_ret_pred(uint32_t saddr, uint32_t v, uint32_t pred) {
  uint32_t old = 0;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}" : "+r"(old) : "r"(saddr), "r"(v), "r"(pred) : "memory");
  return old;
}
__device__ __forceinline__ void smem_sum_i64_pred(uint32_t saddr, int64_t* gslot, uint32_t vl, int32_t vh, uint32_t pred) {
  const uint32_t old = smem_add_ret_pred(saddr, vl, pred);
  const int32_t hi = vh + (int32_t)((uint32_t)(old + vl) < old);
  if (pred && hi != 0) red_add_u64(gslot, (uint64_t)(int64_t)hi << 32);
}

__device__ __forceinline__ void smem_sum_i64(int8_t* tab, int64_t* garr, uint32_t e, uint32_t vl, int32_t vh) {
  const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(tab) + e, vl);
  const int32_t hi = vh + (int32_t)((uint32_t)(old + vl) < old);
  if (hi != 0) red_add_u64(garr + e, (uint64_t)(int64_t)hi << 32);
}

__device__ __forceinline__ void smem_minmax(int op, int8_t* tab, uint32_t e, int64_t v) {
  const bool fp = (op == ACC_MIN_F64) | (op == ACC_MAX_F64);
  if (fp) {
    const double d = __longlong_as_double(v);
    if (d != d) return; /* std::min/std::max never pick up a NaN argument (RuntimeFunctions.cpp:1456-1466) */
    v = b2q_f64_to_ord(v);
  }
  long long* p = reinterpret_cast<long long*>(tab) + e;
  const long long cur = *reinterpret_cast<volatile long long*>(p);
  if ((op == ACC_MIN_I64) | (op == ACC_MIN_F64)) { if (v < cur) atomicMin(p, (long long)v); }
  else { if (v > cur) atomicMax(p, (long long)v); }
}

/* warp-level reductions for the single-group (non-grouped) case */
__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int64_t warp_min_i64(int64_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { const int64_t w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
  return v;
}
__device__ __forceinline__ int64_t warp_max_i64(int64_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { const int64_t w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}

/* ---------------------------------------------------------------------------------------------------------
 * TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP) with mbarrier completion
 * ------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}
/* TMA bulk prefetch of a column slab into L2 (cp.async.bulk.prefetch.L2, SASS UBLKPF.L2): no registers, no shared
 * memory, one instruction per slab — the scan's loads then hit L2 (~300 cycles) instead of HBM (~800) */
__device__ __forceinline__ void tma_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

/* ---------------------------------------------------------------------------------------------------------
 * the scan kernel
 * ------------------------------------------------------------------------------------------------------- */
struct ScanArgs {
  DevProgram prog;
  DevLaunch launch;
  SmemPlan smem;
  const int8_t* smem_image; /* identity image of ONE replica in HBM (MODE_SMEM) */
  int32_t prefetch_distance; /* > 0: TMA bulk-prefetch the column slabs of the chunk this CTA will scan D iterations ahead into L2 */
  int32_t pad_;
  int64_t ndv_bitmap_bytes;  /* estimator query: size of the ACC_NDV bitmap (a power of two) */
};

extern __shared__ __align__(128) int8_t b2q_smem[];

/* one chunk: R rows per thread.  FULL = every row of the chunk exists (no tail masking). */
/* JOIN: 0 = no join level, 1 = INNER, 2 = LEFT (separate instantiations: the plain scan and the INNER probe do not pay
 * for the NULL placeholders of the outer join) */
template <int MODE, bool WAGG, bool KEY32, bool FULL, int BLOCK, int JOIN>
__device__ __forceinline__ void process_chunk(const ScanArgs& A, const int8_t* const* __restrict__ cols, int64_t row0,
                                              int64_t frag_rows, int lane, int8_t* my_tab, uint64_t pol, uint64_t pol_tab) {
  /* BLOCK is a compile-time constant so that the R loads of a column are one base pointer + immediate offsets */
  constexpr int nthr = BLOCK;
  const DevProgram& P = A.prog;
  const DevLaunch& Lh = A.launch;
  uint32_t valid = (1u << R) - 1u;
  if (!FULL) {
    valid = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) valid |= (uint32_t)(row0 + (int64_t)j * nthr < frag_rows) << j;
  }

  /* ---- join level: probe the one-to-one table with the outer key; rows without a match leave `valid`
   * (hash_join_idx[_nullable], GroupByRuntime.cpp:283-311; INNER join).  Columns of the inner table are then read at
   * jidx[] — see load32 / load64. ---- */
  int32_t jidx[JOIN ? R : 1];
  int32_t jval[JOIN ? R : 1]; /* value of the inner column that is packed into the join table (DevJoin::packed_col) */
#define JX(c) ((JOIN && P.col_inner[c]) ? jidx : nullptr), ((JOIN && (c) == P.join.packed_col) ? jval : nullptr), (JOIN == 2 ? P.col_null + (c) : nullptr)
  if (JOIN) {
    const DevJoin& J = P.join;
    const int32_t* __restrict__ buff = Lh.join_buff;
    const int32_t* jsm = A.smem.join_off >= 0 ? reinterpret_cast<const int32_t*>(b2q_smem + A.smem.join_off) : nullptr; /* staged copy */
    const bool packed = J.packed_col >= 0;
    const int32_t packed_null = (JOIN == 2 && packed) ? (int32_t)P.col_null[J.packed_col] : 0;
    const int32_t slot16_null = packed ? (int32_t)P.col_null[J.packed_col] : 0; /* a NULL attribute of a matched row */
    uint32_t matched = 0;
    if (J.fk_width == 8) {
      int64_t k[R];
      load64<!FULL>(k, cols[J.fk_col], row0, nthr, valid, pol);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const uint64_t d = (uint64_t)(k[j] - J.min_key);
        const bool ok = (valid >> j & 1) && d < (uint64_t)J.entry_count && !(J.nullable && k[j] == J.null_val);
        int32_t idx = -1, val = packed_null;
        if (ok && J.slot16) { /* value-only 16-bit slot in shared memory (DevJoin::slot16) */
          const uint32_t s16 = reinterpret_cast<const uint16_t*>(jsm)[d];
          idx = s16 == 0xFFFFu ? -1 : 0;
          val = s16 >= 0xFFFEu ? slot16_null : (int32_t)(J.slot16_min + (int64_t)s16);
        } else if (ok) {
          if (packed) {
            const int2 e2 = jsm ? reinterpret_cast<const int2*>(jsm)[d]
                                : (J.probe_cg ? __ldcg(reinterpret_cast<const int2*>(buff) + d) : __ldg(reinterpret_cast<const int2*>(buff) + d));
            idx = e2.x;
            val = (JOIN == 2 && e2.x < 0) ? packed_null : e2.y;
          } else {
            idx = jsm ? jsm[d] : __ldg(buff + d);
          }
        }
        jidx[JOIN ? j : 0] = idx;
        jval[JOIN ? j : 0] = val;
        matched |= (uint32_t)(idx >= 0) << j;
      }
    } else {
      int32_t k[R];
      load32<!FULL>(k, cols[J.fk_col], J.fk_width, row0, nthr, valid, pol);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const uint64_t d = (uint64_t)((int64_t)k[j] - J.min_key);
        const bool ok = (valid >> j & 1) && d < (uint64_t)J.entry_count && !(J.nullable && (int64_t)k[j] == J.null_val);
        int32_t idx = -1, val = packed_null;
        if (ok && J.slot16) { /* value-only 16-bit slot in shared memory (DevJoin::slot16) */
          const uint32_t s16 = reinterpret_cast<const uint16_t*>(jsm)[d];
          idx = s16 == 0xFFFFu ? -1 : 0;
          val = s16 >= 0xFFFEu ? slot16_null : (int32_t)(J.slot16_min + (int64_t)s16);
        } else if (ok) {
          if (packed) {
            const int2 e2 = jsm ? reinterpret_cast<const int2*>(jsm)[d]
                                : (J.probe_cg ? __ldcg(reinterpret_cast<const int2*>(buff) + d) : __ldg(reinterpret_cast<const int2*>(buff) + d));
            idx = e2.x;
            val = (JOIN == 2 && e2.x < 0) ? packed_null : e2.y;
          } else {
            idx = jsm ? jsm[d] : __ldg(buff + d);
          }
        }
        jidx[JOIN ? j : 0] = idx;
        jval[JOIN ? j : 0] = val;
        matched |= (uint32_t)(idx >= 0) << j;
      }
    }
    if (JOIN != 2) valid &= matched; /* INNER: no match, no row; LEFT: the row stays and its inner columns are NULL */
  }

  /* ---- key column: issued before the filter when the planner expects most sectors to be needed anyway ---- */
  int32_t k32[KEY32 ? R : 1];
  int64_t k64[KEY32 ? 1 : R];
  const bool has_key = !WAGG && P.key.col >= 0;
  const bool eager_key = P.eager_key;
  if (has_key && eager_key) {
    if (KEY32) load32<true>(reinterpret_cast<int32_t(&)[R]>(k32), cols[P.key.col], P.key.width, row0, nthr, valid, pol, JX(P.key.col));
    else load64<true>(reinterpret_cast<int64_t(&)[R]>(k64), cols[P.key.col], row0, nthr, valid, pol, JX(P.key.col));
  }

  uint32_t pass = eval_filter<FULL, JOIN>(P.filter, cols, row0, nthr, valid, pol, P.col_inner, jidx, P.join.packed_col, jval, P.col_null);

  if (has_key && !eager_key) {
    if (KEY32) load32<true>(reinterpret_cast<int32_t(&)[R]>(k32), cols[P.key.col], P.key.width, row0, nthr, pass, pol, JX(P.key.col));
    else load64<true>(reinterpret_cast<int64_t(&)[R]>(k64), cols[P.key.col], row0, nthr, pass, pol, JX(P.key.col));
  }

  /* ---- group index ---- */
  uint32_t e[R];
#pragma unroll
  for (int j = 0; j < R; ++j) e[j] = 0;
  if (!WAGG && !KEY32 && MODE != MODE_BASELINE && P.n_keys > 1) {
    /* multi-column perfect hash: perfect_key_hash (GroupByAndAggregate.cpp:1549-1597) — mixed-radix index over the
     * NULL-translated keys; one warp-uniform pass per GROUP BY column */
    uint32_t bad = 0;
    const uint32_t kmask = P.eager_key ? valid : pass;
    for (int c = 0; c < P.n_keys; ++c) {
      const DevKeyComp& kc = P.keys[c];
      const int64_t mn = kc.min_val, nullv = kc.null_val;
      const uint64_t card = kc.card;
      const uint32_t mult = kc.mult;
      const bool tr = kc.translate_null;
      if (kc.width == 8) {
        int64_t k[R];
        load64<true>(k, cols[kc.col], row0, nthr, kmask, pol, JX(kc.col));
#pragma unroll
        for (int j = 0; j < R; ++j) {
          int64_t d = k[j] - mn;
          if (kc.div_day) { /* DATE: bucketed by day; a value off the day grid has no reconstructible key */
            d = d / 86400;
            if (k[j] % 86400 != 0 && !(tr && k[j] == nullv)) d = -1;
          }
          if (tr) d = (k[j] == nullv) ? (int64_t)card - 1 : d;
          bad |= (uint32_t)((uint64_t)d >= card) << j;
          e[j] += (uint32_t)d * mult;
        }
      } else {
        int32_t k[R];
        load32<true>(k, cols[kc.col], kc.width, row0, nthr, kmask, pol, JX(kc.col));
#pragma unroll
        for (int j = 0; j < R; ++j) {
          int64_t d = (int64_t)k[j] - mn;
          if (tr) d = ((int64_t)k[j] == nullv) ? (int64_t)card - 1 : d;
          bad |= (uint32_t)((uint64_t)d >= card) << j;
          e[j] += (uint32_t)d * mult;
        }
      }
    }
    bad &= pass;
    if (bad) { atomicCAS(Lh.error, 0, B2Q_ERR_KEY_OUT_OF_RANGE); pass &= ~bad; }
  }
  if (has_key) {
    if (MODE != MODE_BASELINE) {
      uint32_t bad = 0;
      const bool tr = P.key.translate_null;
      if (KEY32) {
        const uint32_t mn = (uint32_t)P.key.min_val, n = (uint32_t)P.key.entry_count, nidx = (uint32_t)P.key.null_idx;
        const int32_t nullv = (int32_t)P.key.null_val;
        if (tr) {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            uint32_t idx = (uint32_t)k32[KEY32 ? j : 0] - mn;
            idx = (k32[KEY32 ? j : 0] == nullv) ? nidx : idx;
            bad |= (uint32_t)(idx >= n) << j;
            e[j] = idx;
          }
        } else {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const uint32_t idx = (uint32_t)k32[KEY32 ? j : 0] - mn;
            bad |= (uint32_t)(idx >= n) << j;
            e[j] = idx;
          }
        }
      } else {
        const int64_t mn = P.key.min_val, nullv = P.key.null_val, nidx = P.key.null_idx;
        const uint64_t n = (uint64_t)P.key.entry_count;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          int64_t idx = k64[KEY32 ? 0 : j] - mn;
          if (P.key.div_day) { /* DATE: (key - min) / bucket (get_group_value_fast, GroupByRuntime.cpp:194-209) */
            idx = idx / 86400;
            if (k64[KEY32 ? 0 : j] % 86400 != 0 && !(tr && k64[KEY32 ? 0 : j] == nullv)) idx = -1;
          }
          if (tr) idx = (k64[KEY32 ? 0 : j] == nullv) ? nidx : idx;
          bad |= (uint32_t)((uint64_t)idx >= n) << j;
          e[j] = (uint32_t)idx;
        }
      }
      bad &= pass;
      if (bad) { atomicCAS(Lh.error, 0, B2Q_ERR_KEY_OUT_OF_RANGE); pass &= ~bad; }
    } else {
      const uint32_t n = (uint32_t)P.key.entry_count;
      const uint64_t magic = P.key.hash_magic;
      const int hw = P.key.hash_key_width;
      unsigned long long* keys = reinterpret_cast<unsigned long long*>(Lh.keys);
      /* get_group_value (GroupByRuntime.cpp:25-48): h = MurmurHash3(key) % entry_count, linear probe.
       * Phase 1 puts the home-slot loads of all R rows in flight together (ld.global.cg: L2 is the coherence point,
       * so a concurrent claim by another SM is visible); phase 2 resolves each row, falling back to the probe loop. */
      int64_t key[R];
      uint32_t h[R];
      unsigned long long first[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        key[j] = KEY32 ? (int64_t)k32[KEY32 ? j : 0] : k64[KEY32 ? 0 : j];
        if (key[j] == P.key.null_val) key[j] = P.key.null_logical; /* ENCODING FIXED: physical NULL -> logical NULL */
        h[j] = (uint32_t)__umul64hi(magic * (uint64_t)murmur3_key(key[j], hw), (uint64_t)n); /* == hash % n */
      }
#pragma unroll
      for (int j = 0; j < R; ++j) first[j] = (pass >> j & 1) ? __ldcg(keys + h[j]) : 0ull;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (!(pass >> j & 1)) continue;
        const unsigned long long want = (unsigned long long)key[j];
        uint32_t p = h[j];
        unsigned long long cur = first[j];
        bool found = false;
        for (;;) { /* claim an EMPTY_KEY_64 slot with a 64-bit CAS; a lost race re-examines the same slot */
          if (cur == (unsigned long long)B2Q_I64_MAX) cur = atomicCAS(keys + p, (unsigned long long)B2Q_I64_MAX, want);
          if (cur == (unsigned long long)B2Q_I64_MAX || cur == want) { found = true; break; }
          p = p + 1 == n ? 0 : p + 1;
          if (p == h[j]) break;
          cur = __ldcg(keys + p);
        }
        if (!found) { atomicCAS(Lh.error, 0, B2Q_ERR_OUT_OF_SLOTS); pass &= ~(1u << j); }
        e[j] = p;
      }
      /* the probe loops diverge per lane; without an explicit reconvergence the warp stays split for the rest of the
       * kernel and every following column load is issued lane by lane (ncu: 3.8 active threads per LDG, one 32-B
       * sector per thread, 23x the algorithmic DRAM traffic) */
      __syncwarp();
    }
  }

  const uint32_t arg_mask = P.eager_args ? valid : pass;

  /* ---- fused fast path: COUNT(*) and/or one integer SUM, both updates of a row under one predicate region
   * (ncu, profiles/r1_scan_c2_v3: the per-accumulator loops spend 30 of 70 instructions/row on predicate extraction,
   * BSSY/BRA/BSYNC and address math; fusing them halves that) ---- */
  if (MODE == MODE_SMEM && !WAGG && P.fused) {
    const int ic = P.fused_cnt, is = P.fused_sum;
    const uint32_t cnt32 = ic >= 0 ? smem_u32(my_tab + A.smem.acc_off[ic]) : 0u;
    if (is < 0) {
#pragma unroll
      for (int j = 0; j < R; ++j) if (pass >> j & 1) atomicAdd(reinterpret_cast<uint32_t*>(my_tab + A.smem.acc_off[ic]) + e[j], 1u);
      return;
    }
    const DevAcc& sa = P.accs[is];
    uint32_t* sum_tab = reinterpret_cast<uint32_t*>(my_tab + A.smem.acc_off[is]);
    uint32_t* cnt_tab = reinterpret_cast<uint32_t*>(my_tab + A.smem.acc_off[ic >= 0 ? ic : is]);
    int64_t* gsum = Lh.accs[is];
    (void)cnt32;
    if (sa.op == ACC_SUM_F64) { /* AVG/SUM(double): CAS-loop add on the (warp-private) replica + the count */
      int64_t v[R];
      load64<true>(v, cols[sa.col], row0, nthr, arg_mask, pol, JX(sa.col));
      double* dsum = reinterpret_cast<double*>(sum_tab);
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (pass >> j & 1) {
          if (ic >= 0) atomicAdd(cnt_tab + e[j], 1u);
          atomicAdd(dsum + e[j], __longlong_as_double(v[j]));
        }
    } else if (sa.width == 8) {
      int64_t v[R];
      load64<true>(v, cols[sa.col], row0, nthr, arg_mask, pol, JX(sa.col));
      if (ic >= 0) {
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (pass >> j & 1) {
            atomicAdd(cnt_tab + e[j], 1u);
            const uint32_t vl = (uint32_t)v[j];
            const uint32_t old = atomicAdd(sum_tab + e[j], vl);
            const int32_t hi = (int32_t)(v[j] >> 32) + (int32_t)((uint32_t)(old + vl) < old);
            if (hi != 0) red_add_u64(gsum + e[j], (uint64_t)(int64_t)hi << 32);
          }
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (pass >> j & 1) {
            const uint32_t vl = (uint32_t)v[j];
            const uint32_t old = atomicAdd(sum_tab + e[j], vl);
            const int32_t hi = (int32_t)(v[j] >> 32) + (int32_t)((uint32_t)(old + vl) < old);
            if (hi != 0) red_add_u64(gsum + e[j], (uint64_t)(int64_t)hi << 32);
          }
      }
    } else {
      int32_t v[R];
      load32<true>(v, cols[sa.col], sa.width, row0, nthr, arg_mask, pol, JX(sa.col));
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (pass >> j & 1) {
          if (ic >= 0) atomicAdd(cnt_tab + e[j], 1u);
          const uint32_t vl = (uint32_t)v[j];
          const uint32_t old = atomicAdd(sum_tab + e[j], vl);
          const int32_t hi = (v[j] >> 31) + (int32_t)((uint32_t)(old + vl) < old);
          if (hi != 0) red_add_u64(gsum + e[j], (uint64_t)(int64_t)hi << 32);
        }
    }
    return;
  }

  /* ---- aggregate updates: one warp-uniform dispatch per accumulator per R rows ---- */
  for (int a = 0; a < P.n_accs; ++a) {
    const DevAcc& acc = P.accs[a];
    const int op = acc.op;
    int64_t* garr = Lh.accs[a];
    int8_t* tab = (MODE == MODE_SMEM) ? my_tab + A.smem.acc_off[a] : nullptr;
    /* flags array when THIS accumulator carries the touched flag for the global-table kernels */
    uint8_t* pig = (MODE != MODE_SMEM && P.touch_piggyback == a) ? reinterpret_cast<uint8_t*>(Lh.accs[P.touch_acc]) : nullptr;

    if (op == ACC_NDV) {
      /* estimator query: linear_probabilistic_count (RuntimeFunctions.cpp:2399-2408, cuda_mapd_rt.cu:1300-1308) over
       * the tuple of int64 sub-keys (codegenEstimator); the bitmap lives in HBM/L2 and saturates quickly, so a
       * plain load filters out the bits that are already set before the atomic OR */
      if (WAGG) {
        uint32_t h[R];
#pragma unroll
        for (int j = 0; j < R; ++j) h[j] = 0;
        for (int c = 0; c < P.n_keys; ++c) {
          const DevKeyComp& kc = P.keys[c];
          int64_t k[R];
          if (kc.width == 8) load64<true>(k, cols[kc.col], row0, nthr, pass, pol, JX(kc.col));
          else {
            int32_t t32[R];
            load32<true>(t32, cols[kc.col], kc.width, row0, nthr, pass, pol, JX(kc.col));
#pragma unroll
            for (int j = 0; j < R; ++j) k[j] = t32[j];
          }
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const int64_t v = (kc.translate_null && k[j] == kc.null_val) ? kc.null_logical : k[j];
            h[j] = murmur_block(murmur_block(h[j], (uint32_t)v), (uint32_t)((uint64_t)v >> 32));
          }
        }
        uint32_t* bitmap = reinterpret_cast<uint32_t*>(garr);
        const uint32_t bits_mask = (uint32_t)(A.ndv_bitmap_bytes * 8ull - 1ull); /* the buffer sizes are powers of two */
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (!(pass >> j & 1)) continue;
          const uint32_t bit_pos = murmur3_fmix(h[j], (uint32_t)P.n_keys * 8u) & bits_mask;
          const uint32_t bit = 1u << (bit_pos & 31u);
          uint32_t* w = bitmap + (bit_pos >> 5);
          if (!(__ldcg(w) & bit)) atomicOr(w, bit);
        }
      }
      continue;
    }

    if (op == ACC_COUNT && acc.col < 0) { /* COUNT(*) */
      if (WAGG) {
        const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(pass));
        if (lane == 0 && c) atomicAdd(reinterpret_cast<uint32_t*>(tab), c);
      } else if (MODE == MODE_SMEM) {
        const uint32_t tab32 = smem_u32(tab);
#pragma unroll
        for (int j = 0; j < R; ++j) smem_inc_pred(tab32 + e[j] * 4u, pass >> j & 1);
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) if (pass >> j & 1) global_split_add_touch(garr, pig, e[j], P.key.entry_count, 1u, 0, pol_tab);
      }
      continue;
    }

    if (op == ACC_TOUCH) { /* "a row reached this group": a byte flag, set at most once per thread view */
      if (MODE == MODE_SMEM) {
#pragma unroll
        for (int j = 0; j < R; ++j) if (pass >> j & 1) reinterpret_cast<uint8_t*>(tab)[e[j]] = 1;
      } else if (P.touch_piggyback < 0) {
        uint8_t* flags = reinterpret_cast<uint8_t*>(garr);
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (!(pass >> j & 1)) continue;
          uint32_t w;
          asm volatile("ld.global.ca.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(w) : "l"(flags + e[j]), "l"(pol_tab)); /* a stale 0 only costs a redundant store */
          if (!w) flags[e[j]] = 1;
        }
      }
      continue;
    }

    const bool narrow = acc.width <= 4 && (op == ACC_COUNT || op == ACC_SUM_I64 || op == ACC_MIN_I64 || op == ACC_MAX_I64);
    if (narrow) {
      /* 1/2/4-byte integer argument: 32-bit registers */
      int32_t v[R];
      load32<true>(v, cols[acc.col], acc.width, row0, nthr, arg_mask, pol, JX(acc.col));
      const uint32_t m = not_skipped32(acc, v, pass);
      if (WAGG) {
        if (op == ACC_COUNT) {
          const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m));
          if (lane == 0 && c) atomicAdd(reinterpret_cast<uint32_t*>(tab), c);
        } else if (op == ACC_SUM_I64) {
          int64_t s = 0;
#pragma unroll
          for (int j = 0; j < R; ++j) s += (m >> j & 1) ? (int64_t)v[j] : 0;
          s = warp_sum_i64(s);
          if (lane == 0 && s) smem_sum_i64(tab, garr, 0, (uint32_t)s, (int32_t)(s >> 32));
        } else {
          const bool is_min = op == ACC_MIN_I64;
          int64_t r = is_min ? B2Q_I64_MAX : B2Q_I64_MIN;
#pragma unroll
          for (int j = 0; j < R; ++j) if (m >> j & 1) r = is_min ? min(r, (int64_t)v[j]) : max(r, (int64_t)v[j]);
          r = is_min ? warp_min_i64(r) : warp_max_i64(r);
          if (lane == 0 && r != (is_min ? B2Q_I64_MAX : B2Q_I64_MIN)) smem_minmax(op, tab, 0, r);
        }
      } else if (MODE == MODE_SMEM) {
        const uint32_t tab32 = smem_u32(tab);
        if (op == ACC_COUNT) {
#pragma unroll
          for (int j = 0; j < R; ++j) smem_inc_pred(tab32 + e[j] * 4u, m >> j & 1);
        } else if (op == ACC_SUM_I64) {
#pragma unroll
          for (int j = 0; j < R; ++j) smem_sum_i64_pred(tab32 + e[j] * 4u, garr + e[j], (uint32_t)v[j], v[j] >> 31, m >> j & 1);
        } else {
#pragma unroll
          for (int j = 0; j < R; ++j) if (m >> j & 1) smem_minmax(op, tab, e[j], (int64_t)v[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) if (m >> j & 1) global_update(op, garr, pig, e[j], P.key.entry_count, (int64_t)v[j], pol_tab);
      }
      continue;
    }

    /* 8-byte argument (BIGINT or DOUBLE), or a narrow column feeding a double aggregate (not produced by the planner) */
    int64_t v[R];
    if (acc.width == 8) load64<true>(v, cols[acc.col], row0, nthr, arg_mask, pol, JX(acc.col));
    else {
      int32_t t32[R];
      load32<true>(t32, cols[acc.col], acc.width, row0, nthr, arg_mask, pol, JX(acc.col));
#pragma unroll
      for (int j = 0; j < R; ++j) v[j] = t32[j];
    }
    const uint32_t m = not_skipped64(acc, v, pass);
    if (WAGG) {
      switch (op) {
        case ACC_COUNT: {
          const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m));
          if (lane == 0 && c) atomicAdd(reinterpret_cast<uint32_t*>(tab), c);
          break;
        }
        case ACC_SUM_I64: {
          int64_t s = 0;
#pragma unroll
          for (int j = 0; j < R; ++j) s += (m >> j & 1) ? v[j] : 0;
          s = warp_sum_i64(s);
          if (lane == 0 && s) smem_sum_i64(tab, garr, 0, (uint32_t)s, (int32_t)(s >> 32));
          break;
        }
        case ACC_SUM_F64: {
          double s = 0;
#pragma unroll
          for (int j = 0; j < R; ++j) s += (m >> j & 1) ? __longlong_as_double(v[j]) : 0.0;
          const uint32_t any = __ballot_sync(0xffffffffu, m != 0);
          s = warp_sum_f64(s);
          if (lane == 0 && any) atomicAdd(reinterpret_cast<double*>(tab), s);
          break;
        }
        default: {
          const bool is_min = (op == ACC_MIN_I64) | (op == ACC_MIN_F64);
          const bool fp = (op == ACC_MIN_F64) | (op == ACC_MAX_F64);
          int64_t r = is_min ? B2Q_I64_MAX : B2Q_I64_MIN;
#pragma unroll
          for (int j = 0; j < R; ++j) {
            if (!(m >> j & 1)) continue;
            int64_t x = v[j];
            if (fp) { const double d = __longlong_as_double(x); if (d != d) continue; x = b2q_f64_to_ord(x); }
            r = is_min ? min(r, x) : max(r, x);
          }
          r = is_min ? warp_min_i64(r) : warp_max_i64(r);
          if (lane == 0 && r != (is_min ? B2Q_I64_MAX : B2Q_I64_MIN)) {
            long long* p = reinterpret_cast<long long*>(tab);
            if (is_min) atomicMin(p, (long long)r); else atomicMax(p, (long long)r);
          }
          break;
        }
      }
    } else if (MODE == MODE_SMEM) {
      switch (op) { /* dispatch hoisted out of the row loop */
        case ACC_COUNT: {
          const uint32_t tab32 = smem_u32(tab);
#pragma unroll
          for (int j = 0; j < R; ++j) smem_inc_pred(tab32 + e[j] * 4u, m >> j & 1);
          break;
        }
        case ACC_SUM_I64: {
          const uint32_t tab32 = smem_u32(tab);
#pragma unroll
          for (int j = 0; j < R; ++j) smem_sum_i64_pred(tab32 + e[j] * 4u, garr + e[j], (uint32_t)v[j], (int32_t)(v[j] >> 32), m >> j & 1);
          break;
        }
        case ACC_SUM_F64:
#pragma unroll
          for (int j = 0; j < R; ++j) if (m >> j & 1) atomicAdd(reinterpret_cast<double*>(tab) + e[j], __longlong_as_double(v[j]));
          break;
        default:
#pragma unroll
          for (int j = 0; j < R; ++j) if (m >> j & 1) smem_minmax(op, tab, e[j], v[j]);
          break;
      }
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) if (m >> j & 1) global_update(op, garr, pig, e[j], P.key.entry_count, v[j], pol_tab);
    }
  }
}
#undef JX

template <int MODE, bool WAGG, bool KEY32, int BLOCK, int JOIN>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) b2q_k_scan(const __grid_constant__ ScanArgs A) {
  const DevProgram& P = A.prog;
  const DevLaunch& Lh = A.launch;
  const int tid = threadIdx.x;
  constexpr int nthr = BLOCK;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int64_t chunk_rows = (int64_t)nthr * R;
  __shared__ uint64_t s_bar;
  int8_t* my_tab = nullptr;

  /* a dimension-sized join table is staged into shared memory next to the group table: the probe is then a shared-memory
   * load instead of one random L2 sector per row (profiles/r1_join_knob_sweep.txt: the L2 sector rate is what bounds
   * the join kernels) */
  const bool stage_join = JOIN && A.smem.join_off >= 0;
  if (MODE == MODE_SMEM || stage_join) {
    /* TMA-stage the identity image into every replica of the CTA-private table */
    const uint32_t rb = (uint32_t)A.smem.replica_bytes;
    const uint32_t nrep = MODE == MODE_SMEM ? (uint32_t)A.smem.replicas : 0u;
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&s_bar, rb * nrep + (stage_join ? (uint32_t)A.smem.join_bytes : 0u));
      for (uint32_t r = 0; r < nrep; ++r) {
        uint32_t off = 0;
        while (off < rb) { /* bulk copies of <= 64 KB, 16-byte granularity (replica_bytes is a multiple of 16) */
          const uint32_t n = min(rb - off, 65536u);
          tma_bulk_g2s(b2q_smem + (size_t)r * rb + off, A.smem_image + off, n, &s_bar);
          off += n;
        }
      }
      if (stage_join) {
        const uint32_t jb = (uint32_t)A.smem.join_bytes;
        uint32_t off = 0;
        while (off < jb) {
          const uint32_t n = min(jb - off, 65536u);
          tma_bulk_g2s(b2q_smem + A.smem.join_off + off, reinterpret_cast<const int8_t*>(Lh.join_buff) + off, n, &s_bar);
          off += n;
        }
      }
    }
    mbar_wait(&s_bar, 0);
    if (MODE == MODE_SMEM) my_tab = b2q_smem + (size_t)(warp & (A.smem.replicas - 1)) * rb;
  }

  /* L2 policy for the column stream: it is read exactly once, so mark it evict-first and keep L2 for what is
   * re-used (the HBM/L2-resident group table of the global-table kernels) */
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  uint64_t pol_tab;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_tab));

  /* chunks are visited in increasing order, so the owning fragment is a moving cursor, not a search */
  int frag = 0;
  int64_t frag_first = 0;                                   /* first chunk of `frag` */
  int64_t next_first = __ldg(Lh.frag_chunk_start + 1);      /* first chunk of frag + 1 */
  /* prefetch cursor (thread 0 only) */
  const int pf_dist = A.prefetch_distance;
  int pf_frag = 0;
  int64_t pf_first = 0, pf_next = next_first;
  for (int64_t chunk = blockIdx.x; chunk < Lh.total_chunks; chunk += gridDim.x) {
    if (pf_dist > 0 && tid == 0) {
      const int64_t pc = chunk + (int64_t)pf_dist * gridDim.x;
      if (pc < Lh.total_chunks) {
        while (pc >= pf_next) {
          ++pf_frag;
          pf_first = pf_next;
          pf_next = __ldg(Lh.frag_chunk_start + pf_frag + 1);
        }
        const int64_t prow = (pc - pf_first) * chunk_rows;
        const int64_t pn = min(chunk_rows, __ldg(Lh.frag_rows + pf_frag) - prow);
        const int8_t* const* pcols = Lh.col_ptrs + (size_t)pf_frag * P.n_cols;
        for (int c = 0; c < P.n_cols; ++c) {
          if (!P.col_prefetch[c]) continue;
          const int w = P.col_width[c];
          const uint32_t bytes = (uint32_t)(pn * w) & ~15u;
          if (bytes) tma_prefetch_l2(pcols[c] + prow * w, bytes);
        }
      }
    }
    while (chunk >= next_first) {
      ++frag;
      frag_first = next_first;
      next_first = __ldg(Lh.frag_chunk_start + frag + 1);
    }
    const int64_t frag_rows = __ldg(Lh.frag_rows + frag);
    const int64_t base_row = (chunk - frag_first) * chunk_rows;
    const int8_t* const* __restrict__ cols = Lh.col_ptrs + (size_t)frag * P.n_cols;
    if (base_row + chunk_rows <= frag_rows)
      process_chunk<MODE, WAGG, KEY32, true, BLOCK, JOIN>(A, cols, base_row + tid, frag_rows, lane, my_tab, pol, pol_tab);
    else
      process_chunk<MODE, WAGG, KEY32, false, BLOCK, JOIN>(A, cols, base_row + tid, frag_rows, lane, my_tab, pol, pol_tab);
  }

  if (MODE == MODE_SMEM) {
    /* flush the CTA-private table into the dense HBM table: one RED per (entry, accumulator) that moved */
    __syncthreads();
    const uint32_t rb = (uint32_t)A.smem.replica_bytes;
    const int nrep = A.smem.replicas;
    const int64_t n = P.key.entry_count;
    for (int a = 0; a < P.n_accs; ++a) {
      const int op = P.accs[a].op;
      int64_t* garr = Lh.accs[a];
      const int8_t* base = b2q_smem + A.smem.acc_off[a];
      for (int64_t i = tid; i < n; i += nthr) {
        switch (op) {
          case ACC_NDV: break; /* lives in HBM only */
          case ACC_TOUCH: {
            uint32_t s = 0;
            for (int r = 0; r < nrep; ++r) s |= reinterpret_cast<const uint8_t*>(base + (size_t)r * rb)[i];
            if (s) reinterpret_cast<uint8_t*>(garr)[i] = 1;
            break;
          }
          case ACC_COUNT:
          case ACC_SUM_I64: {
            uint64_t s = 0;
            for (int r = 0; r < nrep; ++r) s += reinterpret_cast<const uint32_t*>(base + (size_t)r * rb)[i];
            if (s) red_add_u64(garr + i, s);
            break;
          }
          case ACC_SUM_F64: {
            double s = 0;
            bool any = false;
            for (int r = 0; r < nrep; ++r) { const double x = reinterpret_cast<const double*>(base + (size_t)r * rb)[i]; any |= (x != 0.0); s += x; }
            if (any) red_add_f64(garr + i, s);
            break;
          }
          case ACC_MIN_I64:
          case ACC_MIN_F64: {
            int64_t s = B2Q_I64_MAX;
            for (int r = 0; r < nrep; ++r) { const int64_t x = reinterpret_cast<const int64_t*>(base + (size_t)r * rb)[i]; s = x < s ? x : s; }
            if (s != B2Q_I64_MAX) red_min_s64(garr + i, s);
            break;
          }
          default: {
            int64_t s = B2Q_I64_MIN;
            for (int r = 0; r < nrep; ++r) { const int64_t x = reinterpret_cast<const int64_t*>(base + (size_t)r * rb)[i]; s = x > s ? x : s; }
            if (s != B2Q_I64_MIN) red_max_s64(garr + i, s);
            break;
          }
        }
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * one-to-one perfect join table (fill_hash_join_buff, JoinHashTable/Runtime/HashJoinRuntime.cpp:120-216, and its
 * init_hash_join_buff): slot[key - min] = inner row index, NULL keys skipped; a slot claimed twice means the join is
 * not one-to-one (the reference then rebuilds a one-to-many table — outside this path)
 * ------------------------------------------------------------------------------------------------------- */
__global__ void b2q_k_join_build(const int8_t* __restrict__ keys, int width, int64_t n_rows, int64_t min_key, int64_t entry_count,
                                 int nullable, int64_t null_val, int32_t* __restrict__ buff, int32_t* __restrict__ error,
                                 const int8_t* __restrict__ packed_vals, int packed_width) {
  /* packed_vals != nullptr: slots are {int32 row, int32 value}; the value of the winning row is written by the thread
   * that claimed the slot (one winner per slot, so no race) */
  const int slot_words = packed_vals ? 2 : 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
    int64_t k;
    switch (width) {
      case 8: k = reinterpret_cast<const int64_t*>(keys)[row]; break;
      case 4: k = reinterpret_cast<const int32_t*>(keys)[row]; break;
      case 2: k = reinterpret_cast<const int16_t*>(keys)[row]; break;
      default: k = reinterpret_cast<const signed char*>(keys)[row]; break;
    }
    if (nullable && k == null_val) continue;
    const uint64_t d = (uint64_t)(k - min_key);
    if (d >= (uint64_t)entry_count) { atomicCAS(error, 0, B2Q_ERR_KEY_OUT_OF_RANGE); continue; }
    if (atomicCAS(buff + d * slot_words, -1, (int32_t)row) != -1) { atomicCAS(error, 0, B2Q_ERR_UNSUPPORTED); continue; }
    if (packed_vals) {
      int32_t v;
      switch (packed_width) {
        case 4: v = reinterpret_cast<const int32_t*>(packed_vals)[row]; break;
        case 2: v = reinterpret_cast<const int16_t*>(packed_vals)[row]; break;
        case -2: v = reinterpret_cast<const uint16_t*>(packed_vals)[row]; break;
        case -1: v = reinterpret_cast<const uint8_t*>(packed_vals)[row]; break;
        default: v = reinterpret_cast<const signed char*>(packed_vals)[row]; break;
      }
      buff[d * 2 + 1] = v;
    }
  }
}

/* one-to-one row table -> value-only uint16 slots (DevJoin::slot16): value - vmin, 0xFFFE = NULL, 0xFFFF = no row */
__global__ void b2q_k_join_slot16(const int32_t* __restrict__ rows, int64_t entry_count, const int8_t* __restrict__ vals, int width,
                                  int64_t null_val, int64_t vmin, uint16_t* __restrict__ out, int32_t* __restrict__ error) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < entry_count; d += stride) {
    const int32_t row = rows[d];
    uint32_t code = 0xFFFFu;
    if (row >= 0) {
      int64_t v;
      switch (width) {
        case 4: v = reinterpret_cast<const int32_t*>(vals)[row]; break;
        case 2: v = reinterpret_cast<const int16_t*>(vals)[row]; break;
        case -2: v = reinterpret_cast<const uint16_t*>(vals)[row]; break;
        case -1: v = reinterpret_cast<const uint8_t*>(vals)[row]; break;
        default: v = reinterpret_cast<const signed char*>(vals)[row]; break;
      }
      if (v == null_val) code = 0xFFFEu;
      else if ((uint64_t)(v - vmin) < 0xFFFEull) code = (uint32_t)(v - vmin);
      else atomicCAS(error, 0, B2Q_ERR_KEY_OUT_OF_RANGE); /* a value outside its chunk-stats range (stale metadata) */
    }
    out[d] = (uint16_t)code;
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * table initialisation (replaces init_group_by_buffer_gpu, GpuInitGroups.cu:124-171): accumulators to the
 * identity of their reduction, baseline keys to EMPTY_KEY_64, plus the one-replica shared-memory image.
 * ------------------------------------------------------------------------------------------------------- */
struct InitArgs {
  int64_t* accs[B2Q_MAX_ACCS];
  int8_t ops[B2Q_MAX_ACCS];
  int32_t n_accs;
  int64_t entry_count;
  int64_t* keys;      /* or nullptr */
  int8_t* smem_image; /* or nullptr */
  SmemPlan smem;
};

__global__ void b2q_k_init(const __grid_constant__ InitArgs A) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.entry_count; i += stride) {
    for (int a = 0; a < A.n_accs; ++a) {
      const int64_t id = b2q_acc_identity(A.ops[a]);
      if (A.ops[a] == ACC_TOUCH) reinterpret_cast<uint8_t*>(A.accs[a])[i] = 0; else A.accs[a][i] = id;
      if (A.smem_image) {
        int8_t* p = A.smem_image + A.smem.acc_off[a];
        if (A.smem.acc_bytes[a] == 1) reinterpret_cast<uint8_t*>(p)[i] = 0;
        else if (A.smem.acc_bytes[a] == 4) reinterpret_cast<uint32_t*>(p)[i] = 0u;
        else reinterpret_cast<int64_t*>(p)[i] = id;
      }
    }
    if (A.keys) A.keys[i] = B2Q_I64_MAX;
  }
}

/* split (lo[n] | hi[n]) accumulator -> plain int64[n] (needed before the NCCL merge and by materialise) */
__global__ void b2q_k_join_split(const int64_t* __restrict__ split, int64_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint32_t* lo = reinterpret_cast<const uint32_t*>(split);
  const int32_t* hi = reinterpret_cast<const int32_t*>(split) + n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (int64_t)(((uint64_t)(uint32_t)hi[i] << 32) + lo[i]);
}

/* ---------------------------------------------------------------------------------------------------------
 * materialise: dense accumulators -> the reference's row-wise output buffer
 * (layout: QueryMemoryDescriptor.cpp:848-955; empty-entry conventions: ResultSetIteration.cpp:2457-2492)
 * ------------------------------------------------------------------------------------------------------- */
struct MatArgs {
  DevLayout layout;
  const int64_t* accs[B2Q_MAX_ACCS];
  const int64_t* keys;
  int8_t* out;
};

__global__ void b2q_k_materialize(const __grid_constant__ MatArgs A) {
  const DevLayout& L = A.layout;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.entry_count; i += stride) {
    int8_t* row = A.out + i * L.row_size;
    bool touched = true;
    int64_t key = 0;
    int64_t mkey_stored[B2Q_MAX_GROUP_COLS], mkey_proj[B2Q_MAX_GROUP_COLS];
    if (L.n_keys > 1) { /* mixed-radix decomposition of the entry index */
      for (int c = 0; c < L.n_keys; ++c) {
        const DevKeyComp& kc = L.keys[c];
        const int64_t comp = (i / kc.mult) % kc.card;
        const bool is_null_comp = kc.translate_null && comp == (int64_t)kc.card - 1;
        mkey_stored[c] = is_null_comp ? kc.null_stored : kc.min_val + comp * kc.step; /* a NULL key is stored translated: max + (bucket ? bucket : 1) */
        mkey_proj[c] = is_null_comp ? kc.null_logical : kc.min_val + comp * kc.step;
      }
      if (L.touched_acc >= 0) touched = reinterpret_cast<const uint8_t*>(A.accs[L.touched_acc])[i] != 0;
    } else if (L.baseline) {
      key = A.keys[i];
      touched = key != B2Q_I64_MAX;
      if (touched && L.key_width == 4) key = (int64_t)(int32_t)key;
    } else {
      key = (i == L.null_idx) ? L.key_null_val : L.key_min + i * L.key_step;
      if (L.touched_acc >= 0) touched = reinterpret_cast<const uint8_t*>(A.accs[L.touched_acc])[i] != 0;
    }
    if (L.has_key_col && L.columnar) {
      /* int64 key columns (initColumnarGroups, QueryMemoryInitializer.cpp:729-735; keys written by
       * get_columnar_group_bin_offset / set_matching_group_value_perfect_hash_columnar / get_group_value_columnar) */
      if (L.n_keys > 1) {
        for (int c = 0; c < L.n_keys; ++c) reinterpret_cast<int64_t*>(A.out + c * L.key_col_stride)[i] = touched ? mkey_stored[c] : B2Q_I64_MAX;
      } else {
        const int64_t stored = (!L.baseline && i == L.null_idx) ? L.key_null_stored : key;
        reinterpret_cast<int64_t*>(A.out)[i] = touched ? stored : B2Q_I64_MAX;
      }
    } else if (L.has_key_col && L.n_keys > 1) {
      for (int c = 0; c < L.n_keys; ++c) reinterpret_cast<int64_t*>(row)[c] = touched ? mkey_stored[c] : B2Q_I64_MAX;
    } else if (L.has_key_col) {
      if (L.key_width == 4) {
        *reinterpret_cast<int32_t*>(row) = touched ? (int32_t)key : 0x7FFFFFFF;
        *reinterpret_cast<int32_t*>(row + 4) = 0;
      } else {
        /* perfect hash stores the TRANSLATED key (NULL -> max+1), GroupByRuntime.cpp:194-209 */
        const int64_t stored = (!L.baseline && i == L.null_idx) ? L.key_null_stored : key;
        *reinterpret_cast<int64_t*>(row) = touched ? stored : B2Q_I64_MAX;
      }
    }
    int64_t vals[B2Q_MAX_SLOTS];
    for (int s = 0; s < L.n_slots; ++s) {
      const DevSlot& sl = L.slots[s];
      int64_t val = sl.init_val;
      if (touched && sl.kind != SLOT_NONE && sl.width != 0) {
        switch (sl.kind) {
          case SLOT_KEY: val = L.n_keys > 1 ? mkey_proj[sl.key_comp] : key; break;
          case SLOT_COUNT: val = A.accs[sl.acc][i]; break;
          default: {
            const int64_t raw = A.accs[sl.acc][i];
            bool is_null = false;
            if (sl.nn >= 0) is_null = A.accs[sl.nn][i] == 0;
            else if (sl.nn == -2) is_null = raw == sl.identity;
            val = is_null ? sl.init_val : (sl.kind == SLOT_VALUE_ORD ? b2q_ord_to_f64(raw) : (sl.scale_day ? raw * 86400 : raw));
            break;
          }
        }
      }
      vals[s] = val;
    }
    /* keyless layouts: an entry whose marker slot still holds its init value IS empty for every reader
     * (ResultSetIteration.cpp:2457-2476); leave it entirely at the init pattern */
    if (L.keyless_marker >= 0 && vals[L.keyless_marker] == L.slots[L.keyless_marker].init_val) {
      for (int s = 0; s < L.n_slots; ++s) vals[s] = L.slots[s].init_val;
    }
    if (L.columnar) {
      for (int s = 0; s < L.n_slots; ++s) {
        const DevSlot& sl = L.slots[s];
        if (sl.kind == SLOT_NONE || sl.width == 0) continue;
        if (sl.width == 4) reinterpret_cast<int32_t*>(A.out + sl.offset)[i] = (int32_t)vals[s];
        else reinterpret_cast<int64_t*>(A.out + sl.offset)[i] = vals[s];
      }
      /* an odd number of 4-byte entries leaves 4 bytes of column padding; the pool buffer is recycled */
      if (i == L.entry_count - 1 && (L.entry_count & 1))
        for (int s = 0; s < L.n_slots; ++s)
          if (L.slots[s].width == 4 && L.slots[s].kind != SLOT_NONE) reinterpret_cast<int32_t*>(A.out + L.slots[s].offset)[L.entry_count] = 0;
      continue;
    }
    int end = 0;
    for (int s = 0; s < L.n_slots; ++s) {
      const DevSlot& sl = L.slots[s];
      if (sl.kind == SLOT_NONE || sl.width == 0) continue;
      if (sl.width == 4) *reinterpret_cast<int32_t*>(row + sl.offset) = (int32_t)vals[s];
      else *reinterpret_cast<int64_t*>(row + sl.offset) = vals[s];
      end = max(end, (int)sl.offset + sl.width);
    }
    /* an odd number of 4-byte slots leaves alignment padding (QueryMemoryDescriptor.cpp:848-860); the buffer comes from
     * a recycled pool, so write the zeros the reference's freshly allocated buffer would hold */
    if (end) for (; end + 4 <= L.row_size; end += 4) *reinterpret_cast<int32_t*>(row + end) = 0;
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * synthetic columns: same counter-based generator as oracle/oracle_gen.h
 * ------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void b2q_k_gen(void* dst, int sql_type, uint64_t seed, uint32_t col_tag, int64_t row0, int64_t count,
                          int64_t lo, uint64_t span, int64_t key_stride) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += step) {
    const uint64_t u = splitmix64(seed ^ ((uint64_t)col_tag << 56) ^ (uint64_t)(row0 + i));
    switch (sql_type) {
      case B2Q_kDOUBLE: static_cast<double*>(dst)[i] = (double)(u >> 11) * (1.0 / 9007199254740992.0); break;
      case B2Q_kBIGINT: static_cast<int64_t*>(dst)[i] = lo + (int64_t)(u % span) * key_stride; break;
      case B2Q_kINT: static_cast<int32_t*>(dst)[i] = (int32_t)(lo + (int64_t)(u % span)); break;
      case B2Q_kSMALLINT: static_cast<int16_t*>(dst)[i] = (int16_t)(lo + (int64_t)(u % span)); break;
      default: static_cast<int8_t*>(dst)[i] = (int8_t)(lo + (int64_t)(u % span)); break;
    }
  }
}

/* =========================================================================================================
 * host-side launch wrappers (called from executor.cpp)
 * ======================================================================================================= */
static int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

struct ScanConfig {
  int block;
  int grid;
  size_t smem_bytes;
};

template <int MODE, bool WAGG, bool KEY32, int BLOCK, int JOIN>
static cudaError_t launch_scan_tb(const ScanArgs& a, const ScanConfig& c, cudaStream_t st) {
  /* the opt-in shared-memory limit is a per-device function attribute: remember which devices have it */
  static unsigned long long attr_set_mask = 0;
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  const bool attr_set = cur_dev < 64 && (attr_set_mask >> cur_dev & 1ull);
  if (!attr_set) {
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, b2q_k_scan<MODE, WAGG, KEY32, BLOCK, JOIN>);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(b2q_k_scan<MODE, WAGG, KEY32, BLOCK, JOIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return e;
    if (cur_dev < 64) attr_set_mask |= 1ull << cur_dev;
  }
  b2q_k_scan<MODE, WAGG, KEY32, BLOCK, JOIN><<<c.grid, c.block, c.smem_bytes, st>>>(a);
  return cudaGetLastError();
}

template <int MODE, bool WAGG, bool KEY32>
static cudaError_t launch_scan_t(const ScanArgs& a, const ScanConfig& c, cudaStream_t st) {
  if (a.prog.join.fk_col >= 0 && a.prog.join.left) /* one hash-join level: separate instantiations, the plain scan stays as it was */
    return c.block == 1024 ? launch_scan_tb<MODE, WAGG, KEY32, 1024, 2>(a, c, st) : launch_scan_tb<MODE, WAGG, KEY32, 512, 2>(a, c, st);
  if (a.prog.join.fk_col >= 0)
    return c.block == 1024 ? launch_scan_tb<MODE, WAGG, KEY32, 1024, 1>(a, c, st) : launch_scan_tb<MODE, WAGG, KEY32, 512, 1>(a, c, st);
  return c.block == 1024 ? launch_scan_tb<MODE, WAGG, KEY32, 1024, 0>(a, c, st) : launch_scan_tb<MODE, WAGG, KEY32, 512, 0>(a, c, st);
}

int scan_rows_per_chunk(int block) { return block * R; }

/* block/grid policy: one CTA per SM with 1024 threads when the table needs more than half of the shared memory,
 * otherwise two CTAs of 512 threads per SM (better tail behaviour, same number of resident threads). */
void scan_config(const B2QQuery& q, int* block, int* ctas_per_sm) {
  const bool big_table = q.smem.total_bytes > 100 * 1024; /* group-table replicas and/or a staged join table */
  *block = big_table ? 1024 : 512;
  *ctas_per_sm = big_table ? 1 : 2;
}

cudaError_t launch_scan(const B2QQuery& q, const DevLaunch& launch, const int8_t* smem_image, int block, int ctas_per_sm,
                        int prefetch_distance, cudaStream_t st) {
  ScanArgs a;
  a.prog = q.prog;
  a.launch = launch;
  a.smem = q.smem;
  a.smem_image = smem_image;
  a.prefetch_distance = prefetch_distance;
  a.pad_ = 0;
  a.ndv_bitmap_bytes = q.plan.query_desc_type == B2Q_Estimator ? q.plan.buffer_size : 0;
  ScanConfig c;
  c.block = block;
  const int64_t max_ctas = (int64_t)sm_count() * ctas_per_sm;
  c.grid = (int)(launch.total_chunks < max_ctas ? (launch.total_chunks > 0 ? launch.total_chunks : 1) : max_ctas);
  c.smem_bytes = (size_t)q.smem.total_bytes; /* 0 for the HBM-table kernels unless a join table is staged */
  const int kernel = q.plan.kernel;
  const bool key32 = q.prog.n_keys <= 1 && q.prog.key.col >= 0 && q.prog.key.width <= 4;
  if (kernel == B2Q_KERNEL_NON_GROUPED) {
    c.smem_bytes = (size_t)q.smem.total_bytes;
    return launch_scan_t<MODE_SMEM, true, false>(a, c, st);
  }
  if (kernel == B2Q_KERNEL_PERFECT_SMEM) {
    c.smem_bytes = (size_t)q.smem.total_bytes;
    return key32 ? launch_scan_t<MODE_SMEM, false, true>(a, c, st) : launch_scan_t<MODE_SMEM, false, false>(a, c, st);
  }
  if (kernel == B2Q_KERNEL_PERFECT_GLOBAL)
    return key32 ? launch_scan_t<MODE_GLOBAL, false, true>(a, c, st) : launch_scan_t<MODE_GLOBAL, false, false>(a, c, st);
  return key32 ? launch_scan_t<MODE_BASELINE, false, true>(a, c, st) : launch_scan_t<MODE_BASELINE, false, false>(a, c, st);
}

cudaError_t launch_join_build(const int8_t* keys, int width, int64_t n_rows, int64_t min_key, int64_t entry_count, int nullable,
                              int64_t null_val, int32_t* buff, int32_t* error, const int8_t* packed_vals, int packed_width,
                              cudaStream_t st) {
  if (entry_count > 0) {
    cudaError_t e = cudaMemsetAsync(buff, 0xFF, (size_t)entry_count * (packed_vals ? 8 : 4), st); /* init_hash_join_buff: every slot -1 */
    if (e != cudaSuccess) return e;
  }
  if (n_rows <= 0 || entry_count <= 0) return cudaSuccess;
  const int block = 256;
  int64_t blocks = (n_rows + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  b2q_k_join_build<<<(int)blocks, block, 0, st>>>(keys, width, n_rows, min_key, entry_count, nullable, null_val, buff, error, packed_vals, packed_width);
  return cudaGetLastError();
}

cudaError_t launch_join_slot16(const int32_t* rows, int64_t entry_count, const int8_t* vals, int width, int64_t null_val, int64_t vmin,
                               uint16_t* out, int32_t* error, cudaStream_t st) {
  if (entry_count <= 0) return cudaSuccess;
  const int block = 256;
  int64_t blocks = (entry_count + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  b2q_k_join_slot16<<<(int)blocks, block, 0, st>>>(rows, entry_count, vals, width, null_val, vmin, out, error);
  return cudaGetLastError();
}

cudaError_t launch_init(const B2QQuery& q, int64_t* const* accs, int64_t* keys, int8_t* smem_image, cudaStream_t st) {
  InitArgs a;
  a.n_accs = q.prog.n_accs;
  for (int i = 0; i < q.prog.n_accs; ++i) { a.accs[i] = accs[i]; a.ops[i] = q.prog.accs[i].op; }
  a.entry_count = q.plan.entry_count;
  a.keys = keys;
  a.smem_image = smem_image;
  a.smem = q.smem;
  const int block = 256;
  int64_t blocks = (q.plan.entry_count + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  b2q_k_init<<<(int)blocks, block, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_join_split(const int64_t* split, int64_t* out, int64_t n, cudaStream_t st) {
  const int block = 256;
  int64_t blocks = (n + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  b2q_k_join_split<<<(int)blocks, block, 0, st>>>(split, out, n);
  return cudaGetLastError();
}

cudaError_t launch_materialize(const B2QQuery& q, const int64_t* const* accs, const int64_t* keys, int8_t* out,
                               cudaStream_t st) {
  MatArgs a;
  a.layout = q.layout;
  for (int i = 0; i < q.prog.n_accs; ++i) a.accs[i] = accs[i];
  a.keys = keys;
  a.out = out;
  const int block = 256;
  int64_t blocks = (q.plan.entry_count + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  b2q_k_materialize<<<(int)blocks, block, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_gen(void* dst, int sql_type, uint64_t seed, uint32_t col_tag, int64_t row0, int64_t count, int64_t lo,
                       int64_t span, int64_t stride, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  const int block = 256;
  int64_t blocks = (count + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  b2q_k_gen<<<(int)blocks, block, 0, st>>>(dst, sql_type, seed, col_tag, row0, count, lo, (uint64_t)(span > 0 ? span : 1), stride ? stride : 1);
  return cudaGetLastError();
}

}  // namespace b2q
