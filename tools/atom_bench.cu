/*
 * atom_bench.cu — what bounds `GROUP BY <1e7 dense keys> SUM(v)` (bench config c4) on a B200?
 * (design experiment behind MODE_GLOBAL of heavydb_b200/csrc/scan_kernel.cuh; profiles/r2_atom_bench.txt holds the output)
 *
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/atom_bench tools/atom_bench.cu && tools/atom_bench [log2 rows]
 *
 * Every variant streams two int64 columns {key in [0, G), v} (evict_first, like the scan kernel) and updates table[key]:
 *   stream    no table update at all: the floor the 16 B/row column stream sets
 *   atom32    atom.add.u32 WITH return on a 4 B/group table (+ carry test): what the scan kernel does today (lo words, 40 MB)
 *   red32     red.add.u32 without return on the same table (no carry possible: only to see what the return trip costs)
 *   red64     red.add.u64 on an 8 B/group table (80 MB), evict_last policy on the table, evict_first on the stream
 *   red64n    the same without a cache policy on the table
 *   atom32x2  atom32 with 16 rows in flight per thread instead of 8
 *   atom64    atom.add.u64 WITH return on the 8 B/group table (what a `touched` flag piggy-backed on the old value would need)
 *   red64fb   red64 + a `touched` BYTE per group: ld.global.ca of the flag, st only when it reads 0 (10 MB of flags)
 *   red64fw   red64 + a `touched` BIT per group: ld.global.ca of the word, atomicOr only when the bit reads 0 (1.25 MB)
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__global__ void k_gen(int64_t* key, int64_t* val, int64_t n, uint32_t G) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t h = mix(i);
    key[i] = (int64_t)(((h >> 32) * (uint64_t)G) >> 32);
    val[i] = (int64_t)(h & 0xFFFFF);
  }
}
__device__ __forceinline__ int64_t ld_stream(const int64_t* p, uint64_t pol) {
  int64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}

constexpr int BLOCK = 512;
enum { V_STREAM, V_ATOM32, V_RED32, V_RED64, V_RED64N, V_ATOM32X2, V_ATOM64, V_RED64FB, V_RED64FW };

template <int V, int R>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) k_agg(const int64_t* __restrict__ key, const int64_t* __restrict__ val, int64_t n,
                                                             void* table, uint32_t G, unsigned long long* sink, uint8_t* flags) {
  uint64_t pol, polt;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(polt));
  uint32_t* t32 = static_cast<uint32_t*>(table);
  unsigned long long* t64 = static_cast<unsigned long long*>(table);
  unsigned long long acc = 0;
  for (int64_t base = (int64_t)blockIdx.x * BLOCK * R; base < n; base += (int64_t)gridDim.x * BLOCK * R) {
    int64_t k[R], v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; k[j] = i < n ? ld_stream(key + i, pol) : -1; }
#pragma unroll
    for (int j = 0; j < R; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; v[j] = i < n ? ld_stream(val + i, pol) : 0; }
    if (V == V_STREAM) {
#pragma unroll
      for (int j = 0; j < R; ++j) acc += (unsigned long long)(k[j] ^ v[j]);
    } else if (V == V_ATOM32 || V == V_ATOM32X2) {
      uint32_t old[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        old[j] = 0;
        if (k[j] >= 0) asm volatile("atom.global.add.L2::cache_hint.u32 %0, [%1], %2, %3;" : "=r"(old[j]) : "l"(t32 + k[j]), "r"((uint32_t)v[j]), "l"(polt) : "memory");
      }
#pragma unroll
      for (int j = 0; j < R; ++j) if (k[j] >= 0 && (uint32_t)(old[j] + (uint32_t)v[j]) < old[j]) atomicAdd(t32 + G + k[j], 1u);
    } else if (V == V_RED32) {
#pragma unroll
      for (int j = 0; j < R; ++j) if (k[j] >= 0) asm volatile("red.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(t32 + k[j]), "r"((uint32_t)v[j]), "l"(polt) : "memory");
    } else if (V == V_RED64) {
#pragma unroll
      for (int j = 0; j < R; ++j) if (k[j] >= 0) asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(t64 + k[j]), "l"(v[j]), "l"(polt) : "memory");
    } else if (V == V_ATOM64) {
      unsigned long long old[R];
#pragma unroll
      for (int j = 0; j < R; ++j) { old[j] = 1; if (k[j] >= 0) old[j] = atomicAdd(t64 + k[j], (unsigned long long)v[j]); }
#pragma unroll
      for (int j = 0; j < R; ++j) if (old[j] == 0) flags[k[j]] = 1;
    } else if (V == V_RED64FB) {
#pragma unroll
      for (int j = 0; j < R; ++j) if (k[j] >= 0) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(t64 + k[j]), "l"(v[j]) : "memory");
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (k[j] < 0) continue;
        uint32_t w;
        asm volatile("ld.global.ca.u8 %0, [%1];" : "=r"(w) : "l"(flags + k[j]));
        if (!w) flags[k[j]] = 1;
      }
    } else if (V == V_RED64FW) {
#pragma unroll
      for (int j = 0; j < R; ++j) if (k[j] >= 0) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(t64 + k[j]), "l"(v[j]) : "memory");
      uint32_t* bits = reinterpret_cast<uint32_t*>(flags);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (k[j] < 0) continue;
        uint32_t w;
        asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(w) : "l"(bits + (k[j] >> 5)));
        if (!(w >> (k[j] & 31) & 1)) atomicOr(bits + (k[j] >> 5), 1u << (k[j] & 31));
      }
    } else if (V == V_RED64N) {
#pragma unroll
      for (int j = 0; j < R; ++j) if (k[j] >= 0) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(t64 + k[j]), "l"(v[j]) : "memory");
    }
  }
  if (V == V_STREAM && acc == 0x1234567) *sink = acc;
}

template <int V, int R>
static void run(const char* name, const int64_t* key, const int64_t* val, int64_t n, void* table, size_t table_bytes, uint32_t G, unsigned long long* sink, int sms) {
  static uint8_t* flags = nullptr;
  if (!flags) CK(cudaMalloc(&flags, (size_t)G + 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaMemsetAsync(table, 0, table_bytes));
    CK(cudaMemsetAsync(flags, 0, (size_t)G + 64));
    CK(cudaEventRecord(e0));
    k_agg<V, R><<<sms * (1024 / BLOCK), BLOCK>>>(key, val, n, table, G, sink, flags);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  printf("%-9s R=%2d  %8.3f ms  %7.1f GB/s of the 16 B/row stream  %6.2f G updates/s\n", name, R, best, n * 16.0 / best / 1e6, n / best / 1e6);
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 29;
  const int64_t n = int64_t(1) << lg;
  const uint32_t G = 10000000u;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t *key, *val;
  void* table;
  unsigned long long* sink;
  const size_t table_bytes = (size_t)G * 8;
  CK(cudaMalloc(&key, n * 8)); CK(cudaMalloc(&val, n * 8)); CK(cudaMalloc(&table, table_bytes)); CK(cudaMalloc(&sink, 8));
  k_gen<<<sms * 4, 512>>>(key, val, n, G);
  CK(cudaDeviceSynchronize());
  printf("rows 2^%d = %lld, groups %u, SMs %d\n", lg, (long long)n, G, sms);
  run<V_STREAM, 8>("stream", key, val, n, table, table_bytes, G, sink, sms);
  run<V_ATOM32, 8>("atom32", key, val, n, table, table_bytes, G, sink, sms);
  run<V_RED32, 8>("red32", key, val, n, table, table_bytes, G, sink, sms);
  run<V_RED64, 8>("red64", key, val, n, table, table_bytes, G, sink, sms);
  run<V_RED64N, 8>("red64n", key, val, n, table, table_bytes, G, sink, sms);
  run<V_ATOM32X2, 16>("atom32x2", key, val, n, table, table_bytes, G, sink, sms);
  run<V_ATOM64, 8>("atom64", key, val, n, table, table_bytes, G, sink, sms);
  run<V_RED64FB, 8>("red64fb", key, val, n, table, table_bytes, G, sink, sms);
  run<V_RED64FW, 8>("red64fw", key, val, n, table, table_bytes, G, sink, sms);
  return 0;
}
