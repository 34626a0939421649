/*
 * scan_inst.cu — instantiations of b2q_k_scan for ONE (join level, table-mode group); the build compiles this file nine
 * times (-DB2Q_SCAN_JOIN={0,1,2} -DB2Q_SCAN_GROUP={0,1,2}) in parallel.
 */
#include "scan_kernel.cuh"

#ifndef B2Q_SCAN_JOIN
#error "compile with -DB2Q_SCAN_JOIN=0|1|2 -DB2Q_SCAN_GROUP=0|1|2"
#endif

namespace b2q {

#define B2Q_CAT_(a, b, c, d) a##b##c##d
#define B2Q_CAT(a, b, c, d) B2Q_CAT_(a, b, c, d)
#define B2Q_THIS_ENTRY B2Q_CAT(launch_scan_j, B2Q_SCAN_JOIN, _g, B2Q_SCAN_GROUP)

template <int MODE, bool WAGG, bool KEY32>
static cudaError_t by_block(const ScanArgs& a, const ScanConfig& c, cudaStream_t st) {
  return c.block == 1024 ? launch_scan_tb<MODE, WAGG, KEY32, 1024, B2Q_SCAN_JOIN>(a, c, st) : launch_scan_tb<MODE, WAGG, KEY32, 512, B2Q_SCAN_JOIN>(a, c, st);
}

cudaError_t B2Q_THIS_ENTRY(const ScanArgs& a, const ScanConfig& c, bool wagg, bool key32, cudaStream_t st) {
#if B2Q_SCAN_GROUP == 0
  if (wagg) return by_block<MODE_SMEM, true, false>(a, c, st);
  return key32 ? by_block<MODE_SMEM, false, true>(a, c, st) : by_block<MODE_SMEM, false, false>(a, c, st);
#elif B2Q_SCAN_GROUP == 1
  (void)wagg;
  return key32 ? by_block<MODE_GLOBAL, false, true>(a, c, st) : by_block<MODE_GLOBAL, false, false>(a, c, st);
#else
  (void)wagg;
  return key32 ? by_block<MODE_BASELINE, false, true>(a, c, st) : by_block<MODE_BASELINE, false, false>(a, c, st);
#endif
}

}  // namespace b2q
