/*
 * scatter_bench.cu — which way of appending 16-byte tuples to ~P x #CTA regions does HBM3e / the B200 L2 like?
 * (design experiment behind pass 1 of heavydb_b200/csrc/radix_agg.cu; profiles/r2_scatter_bench.txt holds the output)
 *
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/scatter_bench tools/scatter_bench.cu && tools/scatter_bench [log2 tuples]
 *
 * Every variant reads n tuples {key, val} coalesced (evict-first), computes part = hash(key) % P and appends the tuple to
 * the region of (part, this CTA); the append cursor is a shared-memory counter.  Reported: ms and GB/s of (read + written) bytes.
 *   copy      coalesced 16 B in, 16 B out: what the memory system gives a streaming kernel of this shape
 *   direct16  one st.global.v2.b64 per tuple, straight to its place (32 different 128-byte lines per warp instruction)
 *   direct16L the same with an L2 evict_last policy on the stores
 *   pair32    tuples of a region are paired in shared memory (double-buffered parking slot) and leave as ONE 32-byte
 *             sector store (st.global.v4.b64): no partially written sector ever reaches L2
 *   tile8kT   tile8k with the runs written by cp.async.bulk (TMA store) and not waited for: the next tile overlaps the drain
 *   tile      CTA-synchronous: a tile of rows is bucketed in shared memory (count, scan, place) and written out with
 *             consecutive lanes on consecutive tuples of a region (runs of ~tile/P tuples)
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__device__ __forceinline__ uint32_t part_of(uint64_t key, uint32_t P) { return (uint32_t)(((mix(key) >> 32) * (uint64_t)P) >> 32); }

__global__ void k_gen(ulonglong2* t, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) t[i] = make_ulonglong2(mix(i) >> 3, (uint64_t)i);
}

__device__ __forceinline__ ulonglong2 ld_stream(const ulonglong2* p, uint64_t pol) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol));
  return v;
}

constexpr int BLOCK = 1024, R = 8;

__global__ void __launch_bounds__(BLOCK, 1) k_copy(const ulonglong2* in, ulonglong2* out, int64_t n) {
  uint64_t pol; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (int64_t base = (int64_t)blockIdx.x * BLOCK * R; base < n; base += (int64_t)gridDim.x * BLOCK * R) {
    ulonglong2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; v[j] = i < n ? ld_stream(in + i, pol) : make_ulonglong2(0, 0); }
#pragma unroll
    for (int j = 0; j < R; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; if (i < n) out[i] = v[j]; }
  }
}

/* MODE 0: direct16, 1: direct16 + evict_last */
template <int MODE>
__global__ void __launch_bounds__(BLOCK, 1) k_direct(const ulonglong2* in, ulonglong2* out, int64_t n, uint32_t P, uint32_t cap, uint32_t* counts) {
  extern __shared__ uint32_t s_cnt[];
  for (int i = threadIdx.x; i < P; i += BLOCK) s_cnt[i] = 0;
  __syncthreads();
  uint64_t pol, polw;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(polw));
  for (int64_t base = (int64_t)blockIdx.x * BLOCK * R; base < n; base += (int64_t)gridDim.x * BLOCK * R) {
    ulonglong2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; v[j] = i < n ? ld_stream(in + i, pol) : make_ulonglong2(~0ull, 0); }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (v[j].x == ~0ull) continue;
      const uint32_t p = part_of(v[j].x, P);
      const uint32_t pos = atomicAdd(s_cnt + p, 1u);
      if (pos >= cap) continue;
      ulonglong2* dst = out + ((uint64_t)p * gridDim.x + blockIdx.x) * cap + pos;
      if (MODE == 0) asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(dst), "l"(v[j].x), "l"(v[j].y) : "memory");
      else asm volatile("st.global.L2::cache_hint.v2.b64 [%0], {%1, %2}, %3;" ::"l"(dst), "l"(v[j].x), "l"(v[j].y), "l"(polw) : "memory");
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += BLOCK) counts[(size_t)i * gridDim.x + blockIdx.x] = min(s_cnt[i], cap);
}

/* pair32: per partition two parking slots; tuple number c of a region: even -> park in slot (c/2)&1, odd -> take its partner
 * from the slot and store both as one 32-byte sector */
struct Park { ulonglong2 t[2]; };
__global__ void __launch_bounds__(BLOCK, 1) k_pair(const ulonglong2* in, ulonglong2* out, int64_t n, uint32_t P, uint32_t cap, uint32_t* counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  Park* park = reinterpret_cast<Park*>(s_raw);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(park + P);
  uint32_t* ready = s_cnt + P;      /* [P][2]: number of the tuple parked in the slot + 1 (0: nothing yet) */
  for (int i = threadIdx.x; i < P; i += BLOCK) { s_cnt[i] = 0; ready[2 * i] = 0; ready[2 * i + 1] = 0; }
  __syncthreads();
  uint64_t pol; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (int64_t base = (int64_t)blockIdx.x * BLOCK * R; base < n; base += (int64_t)gridDim.x * BLOCK * R) {
    ulonglong2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; v[j] = i < n ? ld_stream(in + i, pol) : make_ulonglong2(~0ull, 0); }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (v[j].x == ~0ull) continue;
      const uint32_t p = part_of(v[j].x, P);
      const uint32_t c = atomicAdd(s_cnt + p, 1u);
      if (c >= cap) continue;
      const uint32_t slot = (c >> 1) & 1u;
      volatile uint32_t* rd = ready + 2 * p + slot;
      if ((c & 1u) == 0) {
        /* the slot's previous tenant (tuple c - 4) must have been picked up: its partner sets ready to c - 4 + 2 = c - 2 ... encoded as "free for c" */
        while (*rd != (c >= 4 ? c - 2 : 0u)) {}
        park[p].t[slot] = v[j];
        __threadfence_block();
        *rd = c + 1;
      } else {
        while (*rd != c) {}
        const ulonglong2 a = park[p].t[slot];
        __threadfence_block();
        *rd = c + 1; /* == (c - 1) + 2: free for tuple c + 3's pair */
        ulonglong2* dst = out + ((uint64_t)p * gridDim.x + blockIdx.x) * cap + (c - 1);
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(a.x), "l"(a.y), "l"(v[j].x), "l"(v[j].y) : "memory");
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += BLOCK) {
    const uint32_t c = min(s_cnt[i], cap);
    if (c & 1u) out[((uint64_t)i * gridDim.x + blockIdx.x) * cap + (c - 1)] = park[i].t[((c - 1) >> 1) & 1u];
    counts[(size_t)i * gridDim.x + blockIdx.x] = c;
  }
}

/* tile: bucket TILE = BLOCK * RT rows in shared memory, then write runs */
template <int RT>
__global__ void __launch_bounds__(BLOCK, 1) k_tile(const ulonglong2* in, ulonglong2* out, int64_t n, uint32_t P, uint32_t cap, uint32_t* counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int TILE = BLOCK * RT;
  ulonglong2* tile = reinterpret_cast<ulonglong2*>(s_raw);          /* [TILE] tuples in partition order */
  uint32_t* s_cur = reinterpret_cast<uint32_t*>(tile + TILE);       /* [P] tuples of the region written so far */
  uint32_t* s_cnt = s_cur + P;                                      /* [P] tuples of this tile */
  uint32_t* s_off = s_cnt + P;                                      /* [P + 1] exclusive scan */
  uint16_t* s_pid = reinterpret_cast<uint16_t*>(s_off + P + 1);     /* [TILE] partition of tile slot */
  __shared__ uint32_t s_warp[32];
  for (int i = threadIdx.x; i < P; i += BLOCK) { s_cur[i] = 0; s_cnt[i] = 0; }
  __syncthreads();
  uint64_t pol; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * TILE; base < n; base += (int64_t)gridDim.x * TILE) {
    ulonglong2 v[RT];
    uint32_t pr[RT], rk[RT];
#pragma unroll
    for (int j = 0; j < RT; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; v[j] = i < n ? ld_stream(in + i, pol) : make_ulonglong2(~0ull, 0); }
#pragma unroll
    for (int j = 0; j < RT; ++j) { pr[j] = v[j].x == ~0ull ? 0xFFFFFFFFu : part_of(v[j].x, P); rk[j] = pr[j] != 0xFFFFFFFFu ? atomicAdd(s_cnt + pr[j], 1u) : 0u; }
    __syncthreads();
    /* exclusive scan of s_cnt[0..P) -> s_off: each thread owns ceil(P / BLOCK) consecutive entries */
    const int per = (P + BLOCK - 1) / BLOCK;
    uint32_t loc = 0;
    for (int q = 0; q < per; ++q) { const int i = threadIdx.x * per + q; if (i < (int)P) loc += s_cnt[i]; }
    uint32_t inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) { uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      s_warp[lane] = w; }
    __syncthreads();
    uint32_t run = inc - loc + (warp ? s_warp[warp - 1] : 0u);
    for (int q = 0; q < per; ++q) { const int i = threadIdx.x * per + q; if (i < (int)P) { s_off[i] = run; run += s_cnt[i]; } }
    if (threadIdx.x == BLOCK - 1) s_off[P] = run;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RT; ++j) if (pr[j] != 0xFFFFFFFFu) { const uint32_t s = s_off[pr[j]] + rk[j]; tile[s] = v[j]; s_pid[s] = (uint16_t)pr[j]; }
    __syncthreads();
    const uint32_t total = s_off[P];
    for (uint32_t s = threadIdx.x; s < total; s += BLOCK) {
      const uint32_t p = s_pid[s];
      const uint32_t pos = s_cur[p] + (s - s_off[p]);
      if (pos < cap) out[((uint64_t)p * gridDim.x + blockIdx.x) * cap + pos] = tile[s];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += BLOCK) { s_cur[i] += s_cnt[i]; s_cnt[i] = 0; }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < P; i += BLOCK) counts[(size_t)i * gridDim.x + blockIdx.x] = min(s_cur[i], cap);
}

/* tileT: like tile, but the runs leave with ONE cp.async.bulk (TMA store, shared -> global) per partition instead of one
 * 16-byte store per tuple, and the CTA does not wait for them: the next tile's loads, count + rank and scan run while the stores
 * drain; only `place` (which overwrites the tile) waits for the bulk group to have READ its source. */
template <int RT>
__global__ void __launch_bounds__(BLOCK, 1) k_tile_tma(const ulonglong2* in, ulonglong2* out, int64_t n, uint32_t P, uint32_t cap, uint32_t* counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int TILE = BLOCK * RT;
  ulonglong2* tile = reinterpret_cast<ulonglong2*>(s_raw);
  uint32_t* s_cur = reinterpret_cast<uint32_t*>(tile + TILE);
  uint32_t* s_cnt = s_cur + P;
  uint32_t* s_off = s_cnt + P;
  __shared__ uint32_t s_warp[32];
  for (int i = threadIdx.x; i < P; i += BLOCK) { s_cur[i] = 0; s_cnt[i] = 0; }
  __syncthreads();
  uint64_t pol; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile32 = (uint32_t)__cvta_generic_to_shared(tile);
  for (int64_t base = (int64_t)blockIdx.x * TILE; base < n; base += (int64_t)gridDim.x * TILE) {
    ulonglong2 v[RT];
    uint32_t pr[RT], rk[RT];
#pragma unroll
    for (int j = 0; j < RT; ++j) { const int64_t i = base + threadIdx.x + (int64_t)j * BLOCK; v[j] = i < n ? ld_stream(in + i, pol) : make_ulonglong2(~0ull, 0); }
#pragma unroll
    for (int j = 0; j < RT; ++j) { pr[j] = v[j].x == ~0ull ? 0xFFFFFFFFu : part_of(v[j].x, P); rk[j] = pr[j] != 0xFFFFFFFFu ? atomicAdd(s_cnt + pr[j], 1u) : 0u; }
    __syncthreads();
    const int per = (P + BLOCK - 1) / BLOCK;
    uint32_t loc = 0;
    for (int q = 0; q < per; ++q) { const int i = threadIdx.x * per + q; if (i < (int)P) loc += s_cnt[i]; }
    uint32_t inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) { uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      s_warp[lane] = w; }
    /* the previous tile's bulk stores must have read the tile before `place` overwrites it */
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    uint32_t run = inc - loc + (warp ? s_warp[warp - 1] : 0u);
    for (int q = 0; q < per; ++q) { const int i = threadIdx.x * per + q; if (i < (int)P) { s_off[i] = run; run += s_cnt[i]; } }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RT; ++j) if (pr[j] != 0xFFFFFFFFu) tile[s_off[pr[j]] + rk[j]] = v[j];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy writes of the tile -> visible to the async proxy */
    __syncthreads();
    for (int p = threadIdx.x; p < (int)P; p += BLOCK) {
      const uint32_t cnt = s_cnt[p], pos = s_cur[p];
      const uint32_t ok = pos < cap ? min(cnt, cap - pos) : 0u;
      if (ok) {
        ulonglong2* dst = out + ((uint64_t)p * gridDim.x + blockIdx.x) * cap + pos;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tile32 + s_off[p] * 16u), "r"(ok * 16u) : "memory");
      }
      s_cur[p] = pos + cnt;
      s_cnt[p] = 0;
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    __syncthreads();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  for (int i = threadIdx.x; i < P; i += BLOCK) counts[(size_t)i * gridDim.x + blockIdx.x] = min(s_cur[i], cap);
}

__global__ void k_check(const ulonglong2* out, const uint32_t* counts, uint32_t P, uint32_t ncta, uint32_t cap, unsigned long long* sum, unsigned long long* cnt, unsigned long long* bad) {
  const uint64_t regions = (uint64_t)P * ncta;
  unsigned long long s = 0, c = 0, b = 0;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < regions; r += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t p = (uint32_t)(r / ncta);
    for (uint32_t i = 0; i < counts[r]; ++i) { const ulonglong2 t = out[r * cap + i]; s += t.y; ++c; b += part_of(t.x, P) != p; }
  }
  atomicAdd(sum, s); atomicAdd(cnt, c); atomicAdd(bad, b);
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 28;
  const int64_t n = int64_t(1) << lg;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int ncta = sms;
  ulonglong2 *in, *out;
  uint32_t* counts;
  unsigned long long* res;
  CK(cudaMalloc(&in, n * 16));
  const uint32_t Pmax = 4096;
  const double slack = 1.25;
  CK(cudaMalloc(&out, (size_t)(n * 16 * slack) + (size_t)Pmax * ncta * 64 * 16));
  CK(cudaMalloc(&counts, (size_t)Pmax * ncta * 4));
  CK(cudaMalloc(&res, 24));
  k_gen<<<sms * 8, 256>>>(in, n);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const unsigned long long want_sum = (unsigned long long)n * (unsigned long long)(n - 1) / 2;
  auto report = [&](const char* name, uint32_t P, uint32_t cap, float ms, bool check) {
    unsigned long long h[3] = {0, 0, 0};
    if (check) {
      CK(cudaMemset(res, 0, 24));
      k_check<<<sms * 4, 256>>>(out, counts, P, ncta, cap, res, res + 1, res + 2);
      CK(cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost));
    }
    printf("%-10s P=%4u  %8.3f ms  %7.1f GB/s (read+write)  %s\n", name, P, ms, 2.0 * n * 16 / ms / 1e6,
           !check ? "" : (h[1] == (unsigned long long)n && h[0] == want_sum && h[2] == 0) ? "ok" : "MISMATCH");
    fflush(stdout);
  };
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k_copy<<<ncta, BLOCK>>>(in, out, n);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) report("copy", 0, 0, ms, false);
  }
  const uint32_t Ps[] = {458, 916, 1832, 3664};
  for (uint32_t P : Ps) {
    const double mean = (double)n / ncta / P;
    const uint32_t cap = ((uint32_t)(mean * 1.0 + 8 * sqrt(mean) + 32) + 1) & ~1u;
    if ((double)P * ncta * cap * 16 > (double)(n * 16 * slack) + (double)Pmax * ncta * 64 * 16) { printf("P=%u: scratch too small\n", P); continue; }
    float ms;
#define RUN(name, launch, check)                                               \
    for (int rep = 0; rep < 2; ++rep) {                                        \
      cudaEventRecord(e0); launch; cudaEventRecord(e1);                        \
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());                     \
      cudaEventElapsedTime(&ms, e0, e1);                                       \
      if (rep) report(name, P, cap, ms, check);                                \
    }
    RUN("direct16", (k_direct<0><<<ncta, BLOCK, P * 4>>>(in, out, n, P, cap, counts)), true);
    RUN("direct16L", (k_direct<1><<<ncta, BLOCK, P * 4>>>(in, out, n, P, cap, counts)), true);
    {
      const size_t sm = (size_t)P * (sizeof(Park) + 12);
      CK(cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      RUN("pair32", (k_pair<<<ncta, BLOCK, sm>>>(in, out, n, P, cap, counts)), true);
    }
    {
      const size_t sm4 = (size_t)BLOCK * 4 * 18 + (size_t)P * 12 + 16;
      CK(cudaFuncSetAttribute(k_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      RUN("tile4k", (k_tile<4><<<ncta, BLOCK, sm4>>>(in, out, n, P, cap, counts)), true);
      const size_t sm8 = (size_t)BLOCK * 8 * 18 + (size_t)P * 12 + 16;
      CK(cudaFuncSetAttribute(k_tile<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      if (sm8 <= 220 * 1024) { RUN("tile8k", (k_tile<8><<<ncta, BLOCK, sm8>>>(in, out, n, P, cap, counts)), true); }
      const size_t smT = (size_t)BLOCK * 8 * 16 + (size_t)P * 12 + 16;
      CK(cudaFuncSetAttribute(k_tile_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      if (smT <= 220 * 1024) { RUN("tile8kT", (k_tile_tma<8><<<ncta, BLOCK, smT>>>(in, out, n, P, cap, counts)), true); }
    }
  }
  return 0;
}
