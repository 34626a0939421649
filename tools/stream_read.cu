// Read-only streaming ceiling on this GPU: how fast can ANY kernel pull bytes from HBM (no writes)?
// Used to put the scan kernel's achieved GB/s in context next to MEASURED_PEAKS.json's copy bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_read tools/stream_read.cu && tools/stream_read
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void __launch_bounds__(512) k_read(const int4* __restrict__ p, size_t n, unsigned long long* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
    int4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += (unsigned)v[u].x + (unsigned)v[u].y + (unsigned)v[u].z + (unsigned)v[u].w;
  }
  for (; i < n; i += stride) { int4 v = p[i]; acc += (unsigned)v.x + v.y + v.z + v.w; }
  if (acc == 0x1234567ull) *out = acc;
}

int main() {
  const size_t bytes = (size_t)16 << 30;
  int4* d; unsigned long long* o;
  cudaMalloc(&d, bytes); cudaMalloc(&o, 8);
  cudaMemset(d, 1, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    float best = 1e9;
    for (int it = 0; it < 8; ++it) {
      cudaEventRecord(e0);
      k_read<8><<<sm * ctas, 512>>>(d, bytes / 16, o);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2 && ms < best) best = ms;
    }
    printf("read-only stream: %d CTAs/SM x 512 thr, 16 GiB, best %.3f ms = %.1f GB/s\n", ctas, best, bytes / best / 1e6);
  }
  return 0;
}
