/*
 * oracle_gen.h — counter-based synthetic column generator (TEST INFRASTRUCTURE, see oracle.cpp header).
 *
 * "Data generation must be on-device and counter-based so that the CPU oracle regenerates the
 * identical stream".  The CUDA side (heavydb_b200/csrc/gen.cu, b2q_gen_column) implements the same function;
 * tests/test_gen.py checks them against each other element by element.
 *
 *   u      = splitmix64(seed ^ ((uint64_t)col_tag << 56) ^ (uint64_t)row)
 *   int    : lo + (int64_t)(u % span)                  (span > 0)
 *   double : (u >> 11) * 2^-53   in [0, 1)             (lo/span ignored)
 *
 * Shapes follow the reference's own synthetic benchmark (Benchmarks/synthetic_benchmark/create_table.py:117-137:
 * uniform INT columns in [1,N]).
 */
#ifndef ORACLE_GEN_H
#define ORACLE_GEN_H
#include <stdint.h>

static inline uint64_t oracle_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

static inline uint64_t oracle_gen_u64(uint64_t seed, uint32_t col_tag, int64_t row) {
  return oracle_splitmix64(seed ^ ((uint64_t)col_tag << 56) ^ (uint64_t)row);
}

static inline int64_t oracle_gen_int(uint64_t seed, uint32_t col_tag, int64_t row, int64_t lo, int64_t span) {
  return lo + (int64_t)(oracle_gen_u64(seed, col_tag, row) % (uint64_t)span);
}

static inline double oracle_gen_double(uint64_t seed, uint32_t col_tag, int64_t row) {
  return (double)(oracle_gen_u64(seed, col_tag, row) >> 11) * (1.0 / 9007199254740992.0);
}

#endif
