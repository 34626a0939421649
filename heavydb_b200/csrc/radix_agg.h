/*
 * radix_agg.h — host/device structures of the two-pass radix-partitioned aggregation (radix_agg.cu).  Not part of the ABI.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b2q_internal.h"

#define B2Q_RADIX_MAX_VALS 7   /* distinct aggregate-argument columns a tuple may carry */
#define B2Q_RADIX_RETRY 77001 /* device -> host: a structure of the radix path was too small; re-run with the per-row probe kernel */

namespace b2q {

struct RadixPlan {
  int32_t n_parts;     /* partitions = ceil(entry_count / 2^log_s), by home-slot range */
  int32_t log_s;       /* log2(home slots per partition) == log2(slots of pass 2's private table) */
  int32_t n_vals;      /* value words per tuple */
  int32_t tuple_words; /* 1 + n_vals */
  int8_t val_col[B2Q_RADIX_MAX_VALS + 1];   /* launch column of value word c */
  int8_t val_width[B2Q_RADIX_MAX_VALS + 1]; /* its width code */
  int8_t acc_val[B2Q_MAX_ACCS];             /* accumulator -> value word, -1: no argument */
  int8_t acc_bytes[B2Q_MAX_ACCS];           /* pass 2, bytes per entry in shared memory: 4 (COUNT / integer SUM low word) or 8 */
  int16_t acc_off[B2Q_MAX_ACCS];            /* byte offset per entry of the accumulator's array (8-byte arrays first) */
  int32_t acc_bytes_total;
  int32_t tile;                             /* pass 1 buckets each chunk in shared memory (tuples of <= 2 words) */
};

struct RadixBuffers {
  int64_t* scratch;      /* [n_parts][n_cta1][cap] tuples */
  uint32_t* counts;      /* [n_parts][n_cta1] */
  uint32_t* work_counter;/* pass 2's partition queue */
};

struct RadixArgs {
  DevProgram prog;
  DevLaunch launch;
  int64_t* scratch;
  uint32_t* counts;
  uint32_t* work_counter;
  uint32_t cap;          /* tuples per region */
  uint32_t pad0_;
  int32_t n_parts, log_s, n_cta1;
  int32_t n_vals, tuple_words;
  int8_t val_col[B2Q_RADIX_MAX_VALS + 1];
  int8_t val_width[B2Q_RADIX_MAX_VALS + 1];
  int8_t acc_val[B2Q_MAX_ACCS];
  int8_t acc_bytes[B2Q_MAX_ACCS];
  int16_t acc_off[B2Q_MAX_ACCS];
  int32_t acc_bytes_total;
  int32_t queue_off;              /* pass 2: byte offset of the per-warp queues of parked tuples inside the dynamic shared memory */
  int64_t chunk_begin, chunk_end;
};

bool radix_plan(const B2QQuery& q, RadixPlan* rp);
size_t radix_smem_pass1(const RadixPlan& rp);
size_t radix_smem_pass2(const B2QQuery& q, const RadixPlan& rp, int n_cta1);
void radix_geometry(const B2QQuery& q, const RadixPlan& rp, int64_t chunks, int* n_cta1, uint32_t* cap);
int radix_chunk_rows();
cudaError_t launch_radix(const B2QQuery& q, const RadixPlan& rp, const DevLaunch& launch, const RadixBuffers& buf, int64_t chunk_begin,
                         int64_t chunk_end, int n_cta1, uint32_t cap, cudaStream_t st);

cudaError_t launch_baseline_merge(const B2QQuery& q, const int64_t* src_keys, const int64_t* const* src_accs, int64_t n_src, int64_t skip_begin,
                                  int64_t skip_end, int64_t* keys, int64_t* const* accs, int32_t* error, cudaStream_t st);

}  // namespace b2q
