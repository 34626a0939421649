/* TEST INFRASTRUCTURE — never linked into the product.
 *
 * A host-side reading of the FILTER part of the device program the planner emits (heavydb_b200/csrc/b2q_internal.h:
 * DevFilter / DevTerm), one row at a time, following the semantics the scan kernel implements
 * (kernels.cu: eval_term, eval_term2, eval_filter).  tests/test_filter_lowering.py uses it to check the planner's
 * lowering — range encoding of comparisons, NULL folding, De Morgan push-down, operand ordering on the 4-deep mask stack,
 * constants of days-encoded DATE columns — against the oracle on the CPU, where no kernel can run. */
#include <cstdint>
#include <cstring>
#include <string>

#include "../../heavydb_b200/csrc/b2q_internal.h"

namespace {
int64_t load_int(const int8_t* base, int width, int64_t row) {
  switch (width) {
    case 8: { int64_t v; memcpy(&v, base + row * 8, 8); return v; }
    case 4: { int32_t v; memcpy(&v, base + row * 4, 4); return v; }
    case 2: { int16_t v; memcpy(&v, base + row * 2, 2); return v; }
    case -2: { uint16_t v; memcpy(&v, base + row * 2, 2); return v; }
    case -1: return static_cast<uint8_t>(base[row]);
    default: return base[row];
  }
}
double as_double(int64_t bits) { double d; memcpy(&d, &bits, 8); return d; }

bool eval_term(const DevTerm& t, const int8_t* col, int64_t row) {
  const bool neg = t.negate;
  if (!t.cmp_fp) {
    const int64_t v = load_int(col, t.width, row);
    bool r;
    if (t.width == 8) r = (static_cast<uint64_t>(v) - static_cast<uint64_t>(t.lo) <= t.span) != neg;
    else r = (static_cast<uint32_t>(static_cast<int32_t>(v)) - static_cast<uint32_t>(t.lo) <= static_cast<uint32_t>(t.span)) != neg;
    if (t.null_check) {
      const int64_t nullv = t.width == 8 ? t.null_bits : static_cast<int64_t>(static_cast<int32_t>(t.null_bits));
      if (v == nullv) r = false;
    }
    return r;
  }
  double d;
  bool isnull;
  if (t.col_is_fp) { int64_t bits; memcpy(&bits, col + row * 8, 8); d = as_double(bits); isnull = d == as_double(t.null_bits); }
  else {
    const int64_t v = load_int(col, t.width, row);
    d = static_cast<double>(v);
    isnull = v == (t.width == 8 ? t.null_bits : static_cast<int64_t>(static_cast<int32_t>(t.null_bits)));
  }
  bool r = ((d >= t.flo) & (d <= t.fhi)) != neg;
  if (t.null_check && isnull) r = false;
  return r;
}

bool eval_term2(const DevTerm& t, const int8_t* c1, const int8_t* c2, int64_t row) {
  const int64_t a = t.col_is_fp ? load_int(c1, 8, row) : load_int(c1, t.width, row);
  const int64_t b = t.col2_is_fp ? load_int(c2, 8, row) : load_int(c2, t.width2, row);
  bool isnull, r;
  const int op = t.op2;
  if (t.cmp_fp) {
    double x, y;
    isnull = false;
    if (t.col_is_fp) { x = as_double(a); isnull |= t.nullable1 && x == as_double(t.null_bits); } else { x = static_cast<double>(a); isnull |= t.nullable1 && a == t.null_bits; }
    if (t.col2_is_fp) { y = as_double(b); isnull |= t.nullable2 && y == as_double(t.null_bits2); } else { y = static_cast<double>(b); isnull |= t.nullable2 && b == t.null_bits2; }
    r = op == B2Q_kEQ ? x == y : op == B2Q_kNE ? x != y : op == B2Q_kLT ? x < y : op == B2Q_kGT ? x > y : op == B2Q_kLE ? x <= y : x >= y;
  } else {
    isnull = (t.nullable1 && a == t.null_bits) || (t.nullable2 && b == t.null_bits2);
    r = op == B2Q_kEQ ? a == b : op == B2Q_kNE ? a != b : op == B2Q_kLT ? a < b : op == B2Q_kGT ? a > b : op == B2Q_kLE ? a <= b : a >= b;
  }
  return r && !isnull;
}
}  // namespace

/* 1 / 0 = the row passes / fails the lowered filter; -1 = the program is not one this emulator reads (join level,
 * stack deeper than the device's).  table_cols[c] = chunk of table column c for the fragment, `row` inside it. */
static int32_t eval_filter_impl(const B2QQuery* q, const void* const* table_cols, int64_t row);
extern "C" int32_t b2q_test_eval_filter(const B2QQuery* q, const void* const* table_cols, int64_t row) {
  if (!q || q->prog.join.fk_col >= 0) return -1;
  return eval_filter_impl(q, table_cols, row);
}
/* Join programs: the caller has already done the probe — table_cols[n_outer + c] is inner column c GATHERED at the matching
 * inner row for every outer row (the chunk's NULL sentinel where a LEFT join found no match), so the filter reads one
 * denormalised row exactly as the kernel reads it through the join index. */
extern "C" int32_t b2q_test_eval_filter_joined(const B2QQuery* q, const void* const* table_cols, int64_t row) {
  if (!q || q->prog.join.fk_col < 0) return -1;
  return eval_filter_impl(q, table_cols, row);
}
static int32_t eval_filter_impl(const B2QQuery* q, const void* const* table_cols, int64_t row) {
  const DevFilter& f = q->prog.filter;
  if (f.n_ops == 0) return 1;
  bool st[4];
  int sp = 0;
  for (int i = 0; i < f.n_ops; ++i) {
    const uint32_t op = f.ops[i], kind = op >> 4;
    if (kind == FOP_TERM) {
      if (sp >= 4) return -1;
      const DevTerm& t = f.terms[op & 15];
      const int8_t* c1 = static_cast<const int8_t*>(table_cols[q->col_ids[t.col]]);
      st[sp++] = t.col2 >= 0 ? eval_term2(t, c1, static_cast<const int8_t*>(table_cols[q->col_ids[t.col2]]), row) : eval_term(t, c1, row);
    } else {
      if (sp < 2) return -1;
      const bool b = st[--sp], a = st[--sp];
      st[sp++] = kind == FOP_AND ? (a && b) : (a || b);
    }
  }
  return sp == 1 ? (st[0] ? 1 : 0) : -1;
}
extern "C" int32_t b2q_test_filter_terms(const B2QQuery* q) { return q ? q->prog.filter.n_terms : -1; }

/* The entry a passing row aggregates into under PERFECT hash, read from the lowered key mapping (DevKey / DevKeyComp) with
 * the semantics of process_chunk's "group index" block: idx = key - min (raw chunk value: days for a days-encoded DATE),
 * `/ 86400` for an 8-byte DATE, NULL -> the translated NULL entry; several GROUP BY columns: mixed radix.
 * -1 = out of range (the kernel raises KEY_OUT_OF_RANGE), -2 = not a perfect-hash program. */
extern "C" int64_t b2q_test_group_index(const B2QQuery* q, const void* const* table_cols, int64_t row) {
  if (!q || q->plan.query_desc_type != B2Q_GroupByPerfectHash) return -2;
  const DevProgram& P = q->prog;
  if (P.n_keys > 1) {
    int64_t e = 0;
    for (int c = 0; c < P.n_keys; ++c) {
      const DevKeyComp& kc = P.keys[c];
      const int64_t k = load_int(static_cast<const int8_t*>(table_cols[q->col_ids[kc.col]]), kc.width, row);
      int64_t d = k - kc.min_val;
      if (kc.div_day) { d = d / 86400; if (k % 86400 != 0 && !(kc.translate_null && k == kc.null_val)) d = -1; }
      if (kc.translate_null && k == kc.null_val) d = static_cast<int64_t>(kc.card) - 1;
      if (static_cast<uint64_t>(d) >= kc.card) return -1;
      e += d * static_cast<int64_t>(kc.mult);
    }
    return e;
  }
  const DevKey& K = P.key;
  if (K.col < 0) return -2;
  const int64_t k = load_int(static_cast<const int8_t*>(table_cols[q->col_ids[K.col]]), K.width, row);
  int64_t idx;
  if (K.width != 8) { /* the 32-bit key path: unsigned difference of the low words */
    idx = static_cast<uint32_t>(static_cast<uint32_t>(static_cast<int32_t>(k)) - static_cast<uint32_t>(K.min_val));
    if (K.translate_null && static_cast<int32_t>(k) == static_cast<int32_t>(K.null_val)) idx = K.null_idx;
  } else {
    idx = k - K.min_val;
    if (K.div_day) { idx = idx / 86400; if (k % 86400 != 0 && !(K.translate_null && k == K.null_val)) idx = -1; }
    if (K.translate_null && k == K.null_val) idx = K.null_idx;
  }
  return static_cast<uint64_t>(idx) < static_cast<uint64_t>(K.entry_count) ? idx : -1;
}
