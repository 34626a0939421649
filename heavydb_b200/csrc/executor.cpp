/*
 * executor.cpp — host side of the path and the C ABI of include/b2q.h.
 *
 * Mirrors, for the one path in scope, what sits between Executor::executeWorkUnit and the result set in the
 * reference:
 *   executeWorkUnitImpl / createKernels / launchKernels      QueryEngine/Execute.cpp:2213-2399,3028,3158
 *   fetchChunks (column pointers; H2D when not resident)      QueryEngine/Execute.cpp:3581, ColumnFetcher.cpp:214-288
 *   launchGpuCode (param block, init, launch, copy back)      QueryEngine/QueryExecutionContext.cpp:211-582
 *   per-block / per-device reduction                          Execute.cpp:1696,1772; ResultSetReduction.cpp:203-396
 *     -> here: CTA tables are merged on the device into one dense table; devices merge by all-reduce of that table
 *   ResultSet iteration                                       QueryEngine/ResultSetIteration.cpp:2086-2220,2457-2492
 *
 * There is no CPU execution path in this file: every compute entry needs a CUDA device.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "b2q_internal.h"

namespace b2q {
int32_t make_query(const B2QExecUnit* u, const B2QTableInfo* t, const B2QExecutionOptions* eo, size_t guess,
                   bool has_card, B2QQuery* out, std::string* err);
int scan_rows_per_chunk(int block);
void scan_config(const B2QQuery& q, int* block, int* ctas_per_sm);
cudaError_t launch_scan(const B2QQuery& q, const DevLaunch& launch, const int8_t* smem_image, int block, int ctas_per_sm,
                        cudaStream_t st);
cudaError_t launch_init(const B2QQuery& q, int64_t* const* accs, int64_t* keys, int8_t* smem_image, cudaStream_t st);
cudaError_t launch_materialize(const B2QQuery& q, const int64_t* const* accs, const int64_t* keys, int8_t* out,
                               cudaStream_t st);
cudaError_t launch_gen(void* dst, int sql_type, uint64_t seed, uint32_t col_tag, int64_t row0, int64_t count, int64_t lo,
                       int64_t span, cudaStream_t st);
}  // namespace b2q

using namespace b2q;

static thread_local std::string g_err;

static int32_t set_err(int32_t code, const std::string& m) {
  g_err = m;
  return code;
}
#define CU(call)                                                                                               \
  do {                                                                                                         \
    cudaError_t e__ = (call);                                                                                  \
    if (e__ != cudaSuccess) {                                                                                  \
      cudaGetLastError();                                                                                      \
      return set_err(e__ == cudaErrorMemoryAllocation ? B2Q_ERR_OUT_OF_GPU_MEM : B2Q_ERR_CUDA,                 \
                     std::string(#call) + ": " + cudaGetErrorString(e__));                                     \
    }                                                                                                          \
  } while (0)

static bool have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return false; }
  return n > 0;
}

/* ---------------------------------------------------------------------------------------------------------- */
struct B2QPartial {
  B2QQuery q;
  int device = 0;
  int64_t* accs[B2Q_MAX_ACCS] = {};
  int64_t* keys = nullptr;
  int8_t* smem_image = nullptr;
  int32_t* d_error = nullptr;
  double scan_ms = 0, init_ms = 0, h2d_bytes = 0;
  int64_t launches = 0;
  ~B2QPartial() {
    for (auto& a : accs) if (a) cudaFree(a);
    if (keys) cudaFree(keys);
    if (smem_image) cudaFree(smem_image);
    if (d_error) cudaFree(d_error);
  }
};

struct B2QResultSet {
  B2QQuery q;
  std::vector<int8_t> buf;
  int64_t cursor = 0;
  int64_t cached_rows = -1;
  double scan_ms = 0, init_ms = 0, mat_ms = 0;
};

/* device-side launch tables for one scan launch over a set of (fragment, column pointer) rows */
struct LaunchTables {
  const int8_t** d_cols = nullptr;
  int64_t* d_rows = nullptr;
  int64_t* d_chunk_start = nullptr;
  ~LaunchTables() {
    if (d_cols) cudaFree(d_cols);
    if (d_rows) cudaFree(d_rows);
    if (d_chunk_start) cudaFree(d_chunk_start);
  }
};

static int32_t alloc_partial(B2QPartial& p, cudaStream_t st) {
  const B2QQuery& q = p.q;
  const size_t n = static_cast<size_t>(q.plan.entry_count);
  for (int a = 0; a < q.prog.n_accs; ++a) CU(cudaMalloc(&p.accs[a], std::max<size_t>(n, 1) * 8));
  if (q.plan.kernel == B2Q_KERNEL_BASELINE_GLOBAL) CU(cudaMalloc(&p.keys, std::max<size_t>(n, 1) * 8));
  if (q.smem.use_smem) CU(cudaMalloc(&p.smem_image, std::max<int>(q.smem.replica_bytes, 16)));
  CU(cudaMalloc(&p.d_error, sizeof(int32_t)));
  CU(cudaMemsetAsync(p.d_error, 0, sizeof(int32_t), st));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, st));
  CU(launch_init(q, p.accs, p.keys, p.smem_image, st));
  CU(cudaEventRecord(e1, st));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  p.init_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return B2Q_OK;
}

/* one scan launch over `nf` fragments whose referenced columns are already in device memory */
static int32_t scan_device_fragments(B2QPartial& p, int nf, const std::vector<const int8_t*>& cols /* nf * n_cols */,
                                     const std::vector<int64_t>& rows, cudaStream_t st, bool time_it) {
  const B2QQuery& q = p.q;
  int block, ctas;
  scan_config(q, &block, &ctas);
  const int64_t chunk_rows = scan_rows_per_chunk(block);
  std::vector<int64_t> chunk_start(nf + 1, 0);
  for (int f = 0; f < nf; ++f) chunk_start[f + 1] = chunk_start[f] + (rows[f] + chunk_rows - 1) / chunk_rows;
  if (chunk_start[nf] == 0) return B2Q_OK;
  LaunchTables t;
  CU(cudaMalloc(&t.d_cols, std::max<size_t>(cols.size(), 1) * sizeof(void*)));
  CU(cudaMalloc(&t.d_rows, nf * sizeof(int64_t)));
  CU(cudaMalloc(&t.d_chunk_start, (nf + 1) * sizeof(int64_t)));
  CU(cudaMemcpyAsync(t.d_cols, cols.data(), cols.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(t.d_rows, rows.data(), nf * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(t.d_chunk_start, chunk_start.data(), (nf + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  DevLaunch L;
  memset(&L, 0, sizeof(L));
  L.col_ptrs = reinterpret_cast<const int8_t* const*>(t.d_cols);
  L.frag_rows = t.d_rows;
  L.frag_chunk_start = t.d_chunk_start;
  L.n_frags = nf;
  L.total_chunks = chunk_start[nf];
  for (int a = 0; a < q.prog.n_accs; ++a) L.accs[a] = p.accs[a];
  L.keys = p.keys;
  L.error = p.d_error;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (time_it) {
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, st));
  }
  CU(launch_scan(q, L, p.smem_image, block, ctas, st));
  p.launches += 1;
  if (time_it) {
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    p.scan_ms += ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  } else {
    CU(cudaStreamSynchronize(st)); /* LaunchTables are freed on return */
  }
  return B2Q_OK;
}

/* host-resident table: stream the referenced columns through two staging buffer sets so that the H2D copy of
 * slice k+1 overlaps the scan of slice k (the reference does the H2D in fetchChunks, unpipelined). */
static int32_t scan_host_table(B2QPartial& p, const B2QTableInfo& tbl, cudaStream_t st) {
  const B2QQuery& q = p.q;
  const int nc = q.prog.n_cols;
  const int64_t slice_rows = int64_t(1) << 24; /* 16 Mi rows per slice */
  int widths[B2Q_MAX_COLS];
  size_t bytes_per_row = 0;
  for (int c = 0; c < nc; ++c) {
    const int t = tbl.col_types[q.col_ids[c]].type;
    widths[c] = (t == B2Q_kTINYINT) ? 1 : (t == B2Q_kSMALLINT) ? 2 : (t == B2Q_kINT) ? 4 : 8;
    bytes_per_row += widths[c];
  }
  int64_t max_frag = 0;
  for (int f = 0; f < tbl.num_fragments; ++f) max_frag = std::max<int64_t>(max_frag, tbl.fragments[f].num_tuples);
  const int64_t cap_rows = std::min(slice_rows, std::max<int64_t>(max_frag, 1));
  struct Stage {
    int8_t* buf[B2Q_MAX_COLS] = {};
    cudaEvent_t copied = nullptr, scanned = nullptr;
    bool busy = false;
  } stage[2];
  cudaStream_t copy_st;
  CU(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
  int32_t rc = B2Q_OK;
  auto cleanup = [&]() {
    for (auto& s : stage) {
      for (auto& b : s.buf) if (b) cudaFree(b);
      if (s.copied) cudaEventDestroy(s.copied);
      if (s.scanned) cudaEventDestroy(s.scanned);
    }
    cudaStreamDestroy(copy_st);
  };
  for (auto& s : stage) {
    for (int c = 0; c < nc; ++c)
      if (cudaMalloc(&s.buf[c], static_cast<size_t>(cap_rows) * widths[c] + 16) != cudaSuccess) { cleanup(); cudaGetLastError(); return set_err(B2Q_ERR_OUT_OF_GPU_MEM, "staging buffers"); }
    cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.scanned, cudaEventDisableTiming);
  }
  int block, ctas;
  scan_config(q, &block, &ctas);
  const int64_t chunk_rows = scan_rows_per_chunk(block);
  /* per-slice launch tables live in one device allocation, written once up front */
  struct Slice { int frag; int64_t row0, rows; };
  std::vector<Slice> slices;
  for (int f = 0; f < tbl.num_fragments; ++f)
    for (int64_t r = 0; r < tbl.fragments[f].num_tuples; r += cap_rows)
      slices.push_back({f, r, std::min<int64_t>(cap_rows, tbl.fragments[f].num_tuples - r)});
  const size_t ns = slices.size();
  if (ns == 0) { cleanup(); return B2Q_OK; }
  std::vector<const int8_t*> h_cols(ns * nc);
  std::vector<int64_t> h_rows(ns), h_cs(ns * 2);
  for (size_t i = 0; i < ns; ++i) {
    for (int c = 0; c < nc; ++c) h_cols[i * nc + c] = stage[i & 1].buf[c];
    h_rows[i] = slices[i].rows;
    h_cs[2 * i] = 0;
    h_cs[2 * i + 1] = (slices[i].rows + chunk_rows - 1) / chunk_rows;
  }
  const int8_t** d_cols = nullptr;
  int64_t *d_rows = nullptr, *d_cs = nullptr;
  auto cleanup2 = [&]() { if (d_cols) cudaFree(d_cols); if (d_rows) cudaFree(d_rows); if (d_cs) cudaFree(d_cs); };
  if (cudaMalloc(&d_cols, h_cols.size() * sizeof(void*)) != cudaSuccess || cudaMalloc(&d_rows, ns * 8) != cudaSuccess ||
      cudaMalloc(&d_cs, ns * 16) != cudaSuccess) { cleanup(); cleanup2(); cudaGetLastError(); return set_err(B2Q_ERR_OUT_OF_GPU_MEM, "launch tables"); }
  cudaMemcpyAsync(d_cols, h_cols.data(), h_cols.size() * sizeof(void*), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_rows, h_rows.data(), ns * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_cs, h_cs.data(), ns * 16, cudaMemcpyHostToDevice, st);
  for (size_t i = 0; i < ns && rc == B2Q_OK; ++i) {
    Stage& s = stage[i & 1];
    const Slice& sl = slices[i];
    if (s.busy) cudaStreamWaitEvent(copy_st, s.scanned, 0); /* the scan that used this buffer set is done */
    for (int c = 0; c < nc; ++c) {
      const int8_t* src = static_cast<const int8_t*>(tbl.fragments[sl.frag].col_buffers[q.col_ids[c]]);
      if (!src) { rc = set_err(B2Q_ERR_INVALID_ARGUMENT, "referenced column has a NULL buffer"); break; }
      const size_t nbytes = static_cast<size_t>(sl.rows) * widths[c];
      if (cudaMemcpyAsync(s.buf[c], src + static_cast<size_t>(sl.row0) * widths[c], nbytes, cudaMemcpyHostToDevice, copy_st) != cudaSuccess) {
        rc = set_err(B2Q_ERR_CUDA, std::string("H2D copy: ") + cudaGetErrorString(cudaGetLastError()));
        break;
      }
      p.h2d_bytes += static_cast<double>(nbytes);
    }
    if (rc != B2Q_OK) break;
    cudaEventRecord(s.copied, copy_st);
    cudaStreamWaitEvent(st, s.copied, 0);
    DevLaunch L;
    memset(&L, 0, sizeof(L));
    L.col_ptrs = reinterpret_cast<const int8_t* const*>(d_cols + i * nc);
    L.frag_rows = d_rows + i;
    L.frag_chunk_start = d_cs + 2 * i;
    L.n_frags = 1;
    L.total_chunks = h_cs[2 * i + 1];
    for (int a = 0; a < q.prog.n_accs; ++a) L.accs[a] = p.accs[a];
    L.keys = p.keys;
    L.error = p.d_error;
    cudaError_t e = launch_scan(q, L, p.smem_image, block, ctas, st);
    if (e != cudaSuccess) { rc = set_err(B2Q_ERR_CUDA, std::string("scan launch: ") + cudaGetErrorString(e)); break; }
    p.launches += 1;
    cudaEventRecord(s.scanned, st);
    s.busy = true;
  }
  cudaStreamSynchronize(copy_st);
  cudaStreamSynchronize(st);
  cleanup();
  cleanup2();
  return rc;
}

static int32_t execute_partial_impl(size_t* guess, const B2QTableInfo* tbl, const B2QExecUnit* u,
                                    const B2QCompilationOptions* co, const B2QExecutionOptions* eo, int32_t has_card,
                                    cudaStream_t st, B2QPartial** out) {
  if (!tbl || !u || !co || !eo || !out) return set_err(B2Q_ERR_INVALID_ARGUMENT, "null argument");
  if (co->device_type != B2Q_DEVICE_GPU) return set_err(B2Q_ERR_UNSUPPORTED, "device_type must be GPU: this path has no CPU execution");
  std::unique_ptr<B2QPartial> p(new B2QPartial());
  std::string err;
  const size_t g = guess ? *guess : 0;
  int32_t rc = make_query(u, tbl, eo, g, has_card != 0, &p->q, &err);
  if (rc != B2Q_OK) return set_err(rc, err);
  if (!have_device()) return set_err(B2Q_ERR_NO_DEVICE, "no CUDA device visible; this path has no CPU fallback");
  if (eo->device_ordinal >= 0) CU(cudaSetDevice(eo->device_ordinal));
  CU(cudaGetDevice(&p->device));
  rc = alloc_partial(*p, st);
  if (rc != B2Q_OK) return rc;
  const B2QQuery& q = p->q;
  if (tbl->memory_level == B2Q_GPU_LEVEL) {
    /* multi-fragment launch: one kernel over every fragment handed to this device (Execute.cpp:3075-3101) */
    const int nf = tbl->num_fragments;
    std::vector<const int8_t*> cols(static_cast<size_t>(nf) * q.prog.n_cols);
    std::vector<int64_t> rows(nf);
    for (int f = 0; f < nf; ++f) {
      rows[f] = tbl->fragments[f].num_tuples;
      for (int c = 0; c < q.prog.n_cols; ++c) {
        const void* ptr = tbl->fragments[f].col_buffers[q.col_ids[c]];
        if (!ptr && rows[f] > 0) return set_err(B2Q_ERR_INVALID_ARGUMENT, "referenced column has a NULL buffer");
        cols[static_cast<size_t>(f) * q.prog.n_cols + c] = static_cast<const int8_t*>(ptr);
      }
    }
    if (nf > 0) {
      rc = scan_device_fragments(*p, nf, cols, rows, st, true);
      if (rc != B2Q_OK) return rc;
    }
  } else if (tbl->memory_level == B2Q_CPU_LEVEL) {
    rc = scan_host_table(*p, *tbl, st);
    if (rc != B2Q_OK) return rc;
  } else {
    return set_err(B2Q_ERR_INVALID_ARGUMENT, "memory_level must be B2Q_CPU_LEVEL or B2Q_GPU_LEVEL");
  }
  int32_t dev_err = 0;
  CU(cudaMemcpyAsync(&dev_err, p->d_error, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (dev_err) return set_err(dev_err, dev_err == B2Q_ERR_OUT_OF_SLOTS ? "group-by table is full (OUT_OF_SLOTS)" : "group key outside the chunk-stats range");
  *out = p.release();
  return B2Q_OK;
}

static int32_t finalize_impl(B2QPartial* p, cudaStream_t st, B2QResultSet** out) {
  if (!p || !out) return set_err(B2Q_ERR_INVALID_ARGUMENT, "null argument");
  CU(cudaSetDevice(p->device));
  std::unique_ptr<B2QResultSet> rs(new B2QResultSet());
  rs->q = p->q;
  rs->scan_ms = p->scan_ms;
  rs->init_ms = p->init_ms;
  const size_t nbytes = static_cast<size_t>(p->q.plan.buffer_size);
  rs->buf.resize(nbytes);
  if (nbytes) {
    int8_t* d_out = nullptr;
    CU(cudaMalloc(&d_out, nbytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    cudaError_t e = cudaMemsetAsync(d_out, 0, nbytes, st);
    if (e == cudaSuccess) e = launch_materialize(p->q, p->accs, p->keys, d_out, st);
    cudaEventRecord(e1, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rs->buf.data(), d_out, nbytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    rs->mat_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (e != cudaSuccess) return set_err(B2Q_ERR_CUDA, std::string("materialise: ") + cudaGetErrorString(e));
  }
  *out = rs.release();
  return B2Q_OK;
}

/* ---- result-set iteration ------------------------------------------------------------------------------- */
static bool rs_is_empty_entry(const B2QResultSet* rs, int64_t e) {
  const B2QPlan& p = rs->q.plan;
  if (p.query_desc_type == B2Q_NonGroupedAggregate) return false;
  const int8_t* row = rs->buf.data() + e * p.row_size;
  if (p.keyless_hash) {
    const int s = p.idx_target_as_key;
    int64_t v;
    if (p.slot_padded_width[s] == 4) { int32_t x; memcpy(&x, row + p.slot_offset[s], 4); v = x; }
    else memcpy(&v, row + p.slot_offset[s], 8);
    return v == p.init_vals[s];
  }
  if (p.effective_key_width == 4) { int32_t k; memcpy(&k, row, 4); return k == 0x7FFFFFFF; }
  int64_t k;
  memcpy(&k, row, 8);
  return k == B2Q_I64_MAX;
}

static int type_size(int t) { return t == B2Q_kTINYINT ? 1 : t == B2Q_kSMALLINT ? 2 : t == B2Q_kINT ? 4 : 8; }
static int64_t int_null(int t) { return t == B2Q_kTINYINT ? INT8_MIN : t == B2Q_kSMALLINT ? INT16_MIN : t == B2Q_kINT ? INT32_MIN : INT64_MIN; }

extern "C" {

int32_t b2q_abi_version(void) { return B2Q_ABI_VERSION; }
const char* b2q_last_error_message(void) { return g_err.c_str(); }
const char* b2q_error_string(int32_t code) {
  switch (code) {
    case B2Q_OK: return "No Error";
    case B2Q_ERR_DIV_BY_ZERO: return "Division by zero";
    case B2Q_ERR_OUT_OF_GPU_MEM: return "Query couldn't keep the entire working set of columns in GPU memory";
    case B2Q_ERR_OUT_OF_SLOTS: return "Out of Slots";
    case B2Q_ERR_OVERFLOW_OR_UNDERFLOW: return "Overflow or underflow";
    case B2Q_ERR_OUT_OF_TIME: return "Query execution has exceeded the time limit";
    case B2Q_ERR_INTERRUPTED: return "Query execution has been interrupted";
    case B2Q_ERR_UNSUPPORTED: return "Feature outside the scan/filter/group-by/aggregate path";
    case B2Q_ERR_CARDINALITY_ESTIMATION_REQUIRED: return "CardinalityEstimationRequired";
    case B2Q_ERR_INVALID_ARGUMENT: return "Invalid argument";
    case B2Q_ERR_NO_DEVICE: return "No CUDA device (no CPU fallback on this path)";
    case B2Q_ERR_CUDA: return "CUDA error";
    case B2Q_ERR_KEY_OUT_OF_RANGE: return "Group key outside the chunk-stats range";
    default: return code < 0 ? "Out of Slots (-pos)" : "Unknown error";
  }
}
int32_t b2q_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int32_t b2q_plan(const B2QExecUnit* u, const B2QTableInfo* t, const B2QCompilationOptions* co, const B2QExecutionOptions* eo,
                 size_t guess, int32_t has_card, B2QQuery** out) {
  if (!out || !co) return set_err(B2Q_ERR_INVALID_ARGUMENT, "null argument");
  if (co->device_type != B2Q_DEVICE_GPU) return set_err(B2Q_ERR_UNSUPPORTED, "device_type must be GPU: this path has no CPU execution");
  std::unique_ptr<B2QQuery> q(new B2QQuery());
  std::string err;
  const int32_t rc = make_query(u, t, eo, guess, has_card != 0, q.get(), &err);
  if (rc != B2Q_OK) return set_err(rc, err);
  *out = q.release();
  return B2Q_OK;
}
const B2QPlan* b2q_query_plan(const B2QQuery* q) { return q ? &q->plan : nullptr; }
void b2q_query_free(B2QQuery* q) { delete q; }

int32_t b2q_execute_partial(size_t* guess, int32_t /*is_agg*/, const B2QTableInfo* tbl, const B2QExecUnit* u,
                            const B2QCompilationOptions* co, const B2QExecutionOptions* eo, int32_t has_card,
                            void* stream, B2QPartial** out) {
  return execute_partial_impl(guess, tbl, u, co, eo, has_card, static_cast<cudaStream_t>(stream), out);
}

int32_t b2q_execute_work_unit(size_t* guess, int32_t is_agg, const B2QTableInfo* tbl, const B2QExecUnit* u,
                              const B2QCompilationOptions* co, const B2QExecutionOptions* eo, int32_t has_card,
                              B2QResultSet** out) {
  B2QPartial* p = nullptr;
  int32_t rc = b2q_execute_partial(guess, is_agg, tbl, u, co, eo, has_card, nullptr, &p);
  if (rc != B2Q_OK) return rc;
  rc = finalize_impl(p, nullptr, out);
  delete p;
  return rc;
}

int32_t b2q_partial_num_arrays(const B2QPartial* p) { return p ? p->q.prog.n_accs : 0; }
int32_t b2q_partial_array(const B2QPartial* p, int32_t i, void** ptr, int64_t* count, int32_t* dtype, int32_t* redop) {
  if (!p || i < 0 || i >= p->q.prog.n_accs) return set_err(B2Q_ERR_INVALID_ARGUMENT, "array index");
  const int op = p->q.prog.accs[i].op;
  if (ptr) *ptr = p->accs[i];
  if (count) *count = p->q.plan.entry_count;
  if (dtype) *dtype = op == ACC_SUM_F64 ? B2Q_DT_FLOAT64 : B2Q_DT_INT64;
  if (redop) *redop = (op == ACC_MIN_I64 || op == ACC_MIN_F64) ? B2Q_RED_MIN : (op == ACC_MAX_I64 || op == ACC_MAX_F64) ? B2Q_RED_MAX : B2Q_RED_SUM;
  return B2Q_OK;
}
int32_t b2q_partial_is_mergeable(const B2QPartial* p) { return p && p->q.plan.kernel != B2Q_KERNEL_BASELINE_GLOBAL; }
const B2QPlan* b2q_partial_plan(const B2QPartial* p) { return p ? &p->q.plan : nullptr; }
double b2q_partial_kernel_ms(const B2QPartial* p) { return p ? p->scan_ms : 0; }
int32_t b2q_partial_finalize(B2QPartial* p, void* stream, B2QResultSet** out) { return finalize_impl(p, static_cast<cudaStream_t>(stream), out); }
void b2q_partial_free(B2QPartial* p) { delete p; }

/* Inner entry: same parameter block as the reference's JIT kernel; writes the reference-layout buffer on the
 * device (params->group_by_buffers[0]) instead of returning a host ResultSet. */
int32_t b2q_launch(const B2QQuery* query, const B2QParams* prm, void* stream) {
  if (!query || !prm) return set_err(B2Q_ERR_INVALID_ARGUMENT, "null argument");
  if (!have_device()) return set_err(B2Q_ERR_NO_DEVICE, "no CUDA device visible; this path has no CPU fallback");
  if (prm->join_hash_tables || prm->row_func_mgr) return set_err(B2Q_ERR_UNSUPPORTED, "join hash tables / row function manager");
  if (!prm->num_fragments || !prm->col_buffers || !prm->num_rows || !prm->group_by_buffers)
    return set_err(B2Q_ERR_INVALID_ARGUMENT, "missing kernel parameter");
  if (prm->num_tables && *prm->num_tables != 1) return set_err(B2Q_ERR_UNSUPPORTED, "more than one input table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B2QPartial p;
  p.q = *query;
  if (prm->init_agg_value) {
    for (int s = 0; s < p.q.plan.num_slots; ++s) { p.q.plan.init_vals[s] = prm->init_agg_value[s]; p.q.layout.slots[s].init_val = prm->init_agg_value[s]; }
  }
  CU(cudaGetDevice(&p.device));
  int32_t rc = alloc_partial(p, st);
  if (rc != B2Q_OK) return rc;
  const int nf = static_cast<int>(*prm->num_fragments);
  const int nc = p.q.prog.n_cols;
  std::vector<const int8_t*> cols(static_cast<size_t>(nf) * nc);
  std::vector<int64_t> rows(nf);
  for (int f = 0; f < nf; ++f) {
    rows[f] = prm->num_rows[f];
    for (int c = 0; c < nc; ++c) cols[static_cast<size_t>(f) * nc + c] = prm->col_buffers[f][p.q.col_ids[c]];
  }
  if (nf > 0) {
    rc = scan_device_fragments(p, nf, cols, rows, st, false);
    if (rc != B2Q_OK) return rc;
  }
  int64_t* d_out = nullptr;
  CU(cudaMemcpyAsync(&d_out, prm->group_by_buffers, sizeof(int64_t*), cudaMemcpyDefault, st));
  CU(cudaStreamSynchronize(st));
  CU(launch_materialize(p.q, p.accs, p.keys, reinterpret_cast<int8_t*>(d_out), st));
  int32_t dev_err = 0;
  CU(cudaMemcpyAsync(&dev_err, p.d_error, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (prm->error_codes) cudaMemcpy(prm->error_codes, &dev_err, sizeof(int32_t), cudaMemcpyDefault);
  return dev_err;
}

/* ---- ResultSet surface ---------------------------------------------------------------------------------- */
size_t b2q_rs_entry_count(const B2QResultSet* rs) { return rs ? static_cast<size_t>(rs->q.plan.entry_count) : 0; }
size_t b2q_rs_col_count(const B2QResultSet* rs) { return rs ? static_cast<size_t>(rs->q.plan.num_targets) : 0; }
int32_t b2q_rs_is_row_at_empty(const B2QResultSet* rs, size_t e) { return rs_is_empty_entry(rs, static_cast<int64_t>(e)); }
size_t b2q_rs_row_count(const B2QResultSet* rs) {
  if (!rs) return 0;
  if (rs->cached_rows < 0) {
    int64_t n = 0;
    for (int64_t e = 0; e < rs->q.plan.entry_count; ++e) n += !rs_is_empty_entry(rs, e);
    const_cast<B2QResultSet*>(rs)->cached_rows = n;
  }
  return static_cast<size_t>(rs->cached_rows);
}
int32_t b2q_rs_is_empty(const B2QResultSet* rs) { return b2q_rs_row_count(rs) == 0; }
B2QTypeInfo b2q_rs_get_col_type(const B2QResultSet* rs, size_t col) {
  const B2QTargetInfo& t = rs->q.plan.targets[col];
  if (t.is_agg && t.agg_kind == B2Q_kAVG) return B2QTypeInfo{B2Q_kDOUBLE, 0};
  return t.sql_type;
}
void b2q_rs_move_to_begin(B2QResultSet* rs) { if (rs) rs->cursor = 0; }

int32_t b2q_rs_get_next_row(B2QResultSet* rs, B2QTargetValue* row) {
  const B2QPlan& p = rs->q.plan;
  while (rs->cursor < p.entry_count && rs_is_empty_entry(rs, rs->cursor)) ++rs->cursor;
  if (rs->cursor >= p.entry_count) return 0;
  const int8_t* rowp = rs->buf.data() + rs->cursor * p.row_size;
  ++rs->cursor;
  for (int i = 0; i < p.num_targets; ++i) {
    const B2QTargetInfo& t = p.targets[i];
    const int s = t.first_slot;
    int w = p.slot_padded_width[s];
    const int8_t* ptr = rowp + p.slot_offset[s];
    if (w == 0) { ptr = rowp; w = p.effective_key_width; } /* baseline: the key column is the target */
    int64_t ival;
    if (w == 4) { int32_t x; memcpy(&x, ptr, 4); ival = x; } else memcpy(&ival, ptr, 8);
    B2QTargetValue& o = row[i];
    o.is_fp = 0; o.is_null = 0; o.ival = 0; o.dval = 0;
    /* compact type (get_compact_type): MIN/MAX -> argument type, otherwise the target type */
    const bool has_arg = t.agg_arg_type.type != 0;
    const int compact_type = (t.is_agg && has_arg && (t.agg_kind == B2Q_kMIN || t.agg_kind == B2Q_kMAX)) ? t.agg_arg_type.type : t.sql_type.type;
    if (t.is_agg && t.agg_kind == B2Q_kAVG) { /* pair_to_double, ResultSetBufferAccessors.h:197-227 */
      int64_t cnt;
      memcpy(&cnt, rowp + p.slot_offset[s + 1], 8);
      o.is_fp = 1;
      if (cnt == 0) { o.dval = DBL_MIN; o.is_null = 1; }
      else {
        double dividend;
        if (t.sql_type.type == B2Q_kDOUBLE) memcpy(&dividend, &ival, 8); else dividend = static_cast<double>(ival);
        o.dval = dividend / static_cast<double>(cnt);
        o.is_null = o.dval == DBL_MIN;
      }
      continue;
    }
    if (compact_type == B2Q_kDOUBLE) {
      o.is_fp = 1;
      memcpy(&o.dval, &ival, 8);
      o.is_null = o.dval == DBL_MIN;
      continue;
    }
    int64_t resized = ival;
    switch (type_size(compact_type)) {
      case 1: resized = static_cast<int8_t>(ival); break;
      case 2: resized = static_cast<int16_t>(ival); break;
      case 4: resized = static_cast<int32_t>(ival); break;
      default: break;
    }
    if (resized == int_null(compact_type)) { o.ival = int_null(t.sql_type.type); o.is_null = 1; }
    else o.ival = ival;
  }
  return 1;
}

const int8_t* b2q_rs_storage_buffer(const B2QResultSet* rs, size_t* size_bytes) {
  if (size_bytes) *size_bytes = rs ? rs->buf.size() : 0;
  return rs ? rs->buf.data() : nullptr;
}
const B2QPlan* b2q_rs_query_mem_desc(const B2QResultSet* rs) { return rs ? &rs->q.plan : nullptr; }
double b2q_rs_kernel_ms(const B2QResultSet* rs) { return rs ? rs->scan_ms : 0; }
void b2q_rs_free(B2QResultSet* rs) { delete rs; }

int32_t b2q_gen_column(void* dst, int32_t sql_type, uint64_t seed, uint32_t col_tag, int64_t row0, int64_t count,
                       int64_t lo, int64_t span, void* stream) {
  if (!have_device()) return set_err(B2Q_ERR_NO_DEVICE, "no CUDA device visible");
  CU(launch_gen(dst, sql_type, seed, col_tag, row0, count, lo, span, static_cast<cudaStream_t>(stream)));
  return B2Q_OK;
}

} /* extern "C" */
